// Resampling kernels (NHWC fp32, HBM/L2-bound gathers; backward passes are atomics-free gathers => deterministic).
// Reference call sites: nn.Upsample(scale_factor=2, mode='bicubic', align_corners=False) at
// /root/reference/models/generator.py:221,225; F.interpolate(size=..., mode='bilinear') at generator.py:244;
// F.interpolate(scale_factor=0.5|0.25, mode='bicubic') at GAN_DANet_train.ipynb:226,231; MaxPool2d(2,2) of
// torchvision vgg19.features used by models/losses.py:58.
// Semantics: source index (dst+0.5)*ratio-0.5, cubic A=-0.75 with border-clamped taps,
// bilinear source clamped at 0.
#include <cuda_bf16.h>
#include "common.cuh"

namespace gdn {

__device__ __forceinline__ void cubic_coeffs(float t, float w[4]) {
  const float A = -0.75f;
  float u = t + 1.f;  w[0] = ((A * u - 5.f * A) * u + 8.f * A) * u - 4.f * A;
  u = t;              w[1] = ((A + 2.f) * u - (A + 3.f)) * u * u + 1.f;
  u = 1.f - t;        w[2] = ((A + 2.f) * u - (A + 3.f)) * u * u + 1.f;
  u = 2.f - t;        w[3] = ((A * u - 5.f * A) * u + 8.f * A) * u - 4.f * A;
}
__device__ __forceinline__ int clampi(int v, int lo, int hi) { return v < lo ? lo : (v > hi ? hi : v); }

// source position of output o for ratio = in/out (or 1/scale_factor)
__device__ __forceinline__ void cubic_src(int o, float ratio, int& i0, float& t) {
  float s = (o + 0.5f) * ratio - 0.5f;
  float f = floorf(s);
  i0 = (int)f; t = s - f;
}
__device__ __forceinline__ void linear_src(int o, float ratio, int n_in, int& i0, int& i1, float& t) {
  float s = (o + 0.5f) * ratio - 0.5f;
  if (s < 0.f) s = 0.f;
  i0 = (int)s; if (i0 > n_in - 1) i0 = n_in - 1;
  i1 = i0 + 1 < n_in ? i0 + 1 : n_in - 1;
  t = s - (float)i0;
}

// ------------------------------------------------------------------ bicubic x2 up-sampling
template <int VEC>
__global__ void __launch_bounds__(256) bicubic_up2_fwd_kernel(const float* __restrict__ x, float* __restrict__ y, int B, int H, int W, int C) {
  const int Cv = C / VEC, Ho = 2 * H, Wo = 2 * W;
  const long long total = (long long)B * Ho * Wo * Cv;
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
    int c = (int)(idx % Cv) * VEC; long long r = idx / Cv;
    int ox = (int)(r % Wo); r /= Wo; int oy = (int)(r % Ho); int b = (int)(r / Ho);
    int y0, x0; float ty, tx, wy[4], wx[4];
    cubic_src(oy, 0.5f, y0, ty); cubic_src(ox, 0.5f, x0, tx);
    cubic_coeffs(ty, wy); cubic_coeffs(tx, wx);
    float acc[VEC];
#pragma unroll
    for (int v = 0; v < VEC; ++v) acc[v] = 0.f;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      int iy = clampi(y0 - 1 + i, 0, H - 1);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        int ix = clampi(x0 - 1 + j, 0, W - 1);
        const float* p = x + (((size_t)b * H + iy) * W + ix) * C + c;
        float wgt = wy[i] * wx[j];
        if (VEC == 4) { float4 q = *reinterpret_cast<const float4*>(p); acc[0] = fmaf(wgt, q.x, acc[0]); acc[1] = fmaf(wgt, q.y, acc[1]); acc[2] = fmaf(wgt, q.z, acc[2]); acc[3] = fmaf(wgt, q.w, acc[3]); }
        else acc[0] = fmaf(wgt, *p, acc[0]);
      }
    }
    float* o = y + (((size_t)b * Ho + oy) * Wo + ox) * C + c;
    if (VEC == 4) *reinterpret_cast<float4*>(o) = make_float4(acc[0], acc[1], acc[2], acc[3]);
    else *o = acc[0];
  }
}

// total weight with which output o (of 2n outputs) reads input i (border clamping folds several taps onto one input)
__device__ __forceinline__ float cubic_up2_weight(int o, int i, int n) {
  int i0; float t, w[4];
  cubic_src(o, 0.5f, i0, t); cubic_coeffs(t, w);
  float s = 0.f;
#pragma unroll
  for (int k = 0; k < 4; ++k) if (clampi(i0 - 1 + k, 0, n - 1) == i) s += w[k];
  return s;
}

template <int VEC>
__global__ void __launch_bounds__(256) bicubic_up2_bwd_kernel(const float* __restrict__ dy, float* __restrict__ dx, int B, int H, int W, int C) {
  const int Cv = C / VEC, Ho = 2 * H, Wo = 2 * W;
  const long long total = (long long)B * H * W * Cv;
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
    int c = (int)(idx % Cv) * VEC; long long r = idx / Cv;
    int ix = (int)(r % W); r /= W; int iy = (int)(r % H); int b = (int)(r / H);
    float wy[8], wx[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      int oy = 2 * iy - 3 + k, ox = 2 * ix - 3 + k;
      wy[k] = (oy >= 0 && oy < Ho) ? cubic_up2_weight(oy, iy, H) : 0.f;
      wx[k] = (ox >= 0 && ox < Wo) ? cubic_up2_weight(ox, ix, W) : 0.f;
    }
    float acc[VEC];
#pragma unroll
    for (int v = 0; v < VEC; ++v) acc[v] = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      if (wy[i] == 0.f) continue;
      int oy = 2 * iy - 3 + i;
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        if (wx[j] == 0.f) continue;
        int ox = 2 * ix - 3 + j;
        const float* p = dy + (((size_t)b * Ho + oy) * Wo + ox) * C + c;
        float wgt = wy[i] * wx[j];
        if (VEC == 4) { float4 q = *reinterpret_cast<const float4*>(p); acc[0] = fmaf(wgt, q.x, acc[0]); acc[1] = fmaf(wgt, q.y, acc[1]); acc[2] = fmaf(wgt, q.z, acc[2]); acc[3] = fmaf(wgt, q.w, acc[3]); }
        else acc[0] = fmaf(wgt, *p, acc[0]);
      }
    }
    float* o = dx + (((size_t)b * H + iy) * W + ix) * C + c;
    if (VEC == 4) *reinterpret_cast<float4*>(o) = make_float4(acc[0], acc[1], acc[2], acc[3]);
    else *o = acc[0];
  }
}


// ---- 128-bit fast paths ------------------------------------------------------------------------------------------------
// x2 bicubic: output rows 2i / 2i+1 read input rows i-2..i+1 (t = 0.75) / i-1..i+2 (t = 0.25); one thread produces the 2x2
// output block of input pixel (i, j) from the 5x5 clamped neighbourhood, separably (25 loads for 4 outputs instead of 64).
// ADD: y = up2(x) + bilinear_resize(s -> 2H x 2W) in the same pass (the skip fusion of generator.py:242-246 lands on the last up-sampling's output:
// one write of the 64-channel full-resolution tensor instead of write + read-modify-write).
template <int ADD>
__global__ void __launch_bounds__(256) bicubic_up2_fwd_v4_kernel(const float* __restrict__ x, float* __restrict__ y, int B, int H, int W, int C,
                                                                  const float* __restrict__ s, int Hs, int Ws) {
  const int Cv = C >> 2, Wo = 2 * W;
  const long long total = (long long)B * H * W * Cv;
  float w75[4], w25[4];
  cubic_coeffs(0.75f, w75); cubic_coeffs(0.25f, w25);
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(idx % Cv) << 2; long long r = idx / Cv;
    const int j = (int)(r % W); r /= W; const int i = (int)(r % H); const int b = (int)(r / H);
    float4 o[2][2];
#pragma unroll
    for (int a = 0; a < 2; ++a)
#pragma unroll
      for (int e = 0; e < 2; ++e) o[a][e] = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int k = 0; k < 5; ++k) {
      const float* row = x + (((size_t)b * H + clampi(i - 2 + k, 0, H - 1)) * W) * C + c;
      float4 v[5];
#pragma unroll
      for (int m = 0; m < 5; ++m) v[m] = *reinterpret_cast<const float4*>(row + (size_t)clampi(j - 2 + m, 0, W - 1) * C);
      float4 h0 = make_float4(0.f, 0.f, 0.f, 0.f), h1 = h0;
#pragma unroll
      for (int m = 0; m < 4; ++m) {
        h0.x = fmaf(w75[m], v[m].x, h0.x); h0.y = fmaf(w75[m], v[m].y, h0.y); h0.z = fmaf(w75[m], v[m].z, h0.z); h0.w = fmaf(w75[m], v[m].w, h0.w);
        h1.x = fmaf(w25[m], v[m + 1].x, h1.x); h1.y = fmaf(w25[m], v[m + 1].y, h1.y); h1.z = fmaf(w25[m], v[m + 1].z, h1.z); h1.w = fmaf(w25[m], v[m + 1].w, h1.w);
      }
      if (k < 4) {
        const float wv = w75[k];
        o[0][0].x = fmaf(wv, h0.x, o[0][0].x); o[0][0].y = fmaf(wv, h0.y, o[0][0].y); o[0][0].z = fmaf(wv, h0.z, o[0][0].z); o[0][0].w = fmaf(wv, h0.w, o[0][0].w);
        o[0][1].x = fmaf(wv, h1.x, o[0][1].x); o[0][1].y = fmaf(wv, h1.y, o[0][1].y); o[0][1].z = fmaf(wv, h1.z, o[0][1].z); o[0][1].w = fmaf(wv, h1.w, o[0][1].w);
      }
      if (k >= 1) {
        const float wv = w25[k - 1];
        o[1][0].x = fmaf(wv, h0.x, o[1][0].x); o[1][0].y = fmaf(wv, h0.y, o[1][0].y); o[1][0].z = fmaf(wv, h0.z, o[1][0].z); o[1][0].w = fmaf(wv, h0.w, o[1][0].w);
        o[1][1].x = fmaf(wv, h1.x, o[1][1].x); o[1][1].y = fmaf(wv, h1.y, o[1][1].y); o[1][1].z = fmaf(wv, h1.z, o[1][1].z); o[1][1].w = fmaf(wv, h1.w, o[1][1].w);
      }
    }
    if (ADD) {
      const float ry = (float)Hs / (float)(2 * H), rx = (float)Ws / (float)Wo;
      const float* sb = s + (size_t)b * Hs * Ws * C + c;
#pragma unroll
      for (int a = 0; a < 2; ++a) {
        int y0, y1; float ty;
        linear_src(2 * i + a, ry, Hs, y0, y1, ty);
#pragma unroll
        for (int e = 0; e < 2; ++e) {
          int x0, x1; float tx;
          linear_src(2 * j + e, rx, Ws, x0, x1, tx);
          const float w00 = (1.f - ty) * (1.f - tx), w01 = (1.f - ty) * tx, w10 = ty * (1.f - tx), w11 = ty * tx;
          const float4 q00 = __ldg(reinterpret_cast<const float4*>(sb + ((size_t)y0 * Ws + x0) * C)), q01 = __ldg(reinterpret_cast<const float4*>(sb + ((size_t)y0 * Ws + x1) * C));
          const float4 q10 = __ldg(reinterpret_cast<const float4*>(sb + ((size_t)y1 * Ws + x0) * C)), q11 = __ldg(reinterpret_cast<const float4*>(sb + ((size_t)y1 * Ws + x1) * C));
          float4 rr;
          rr.x = w00 * q00.x + w01 * q01.x + w10 * q10.x + w11 * q11.x; rr.y = w00 * q00.y + w01 * q01.y + w10 * q10.y + w11 * q11.y;
          rr.z = w00 * q00.z + w01 * q01.z + w10 * q10.z + w11 * q11.z; rr.w = w00 * q00.w + w01 * q01.w + w10 * q10.w + w11 * q11.w;
          o[a][e].x += rr.x; o[a][e].y += rr.y; o[a][e].z += rr.z; o[a][e].w += rr.w;
        }
      }
    }
#pragma unroll
    for (int a = 0; a < 2; ++a)
#pragma unroll
      for (int e = 0; e < 2; ++e)
        *reinterpret_cast<float4*>(y + (((size_t)b * 2 * H + 2 * i + a) * Wo + 2 * j + e) * C + c) = o[a][e];
  }
}

// weights with which the 8 outputs 2i-3 .. 2i+4 read input i (interior: constants; borders: clamped taps fold together)
__device__ __forceinline__ void up2_bwd_weights(int i, int n, const float w75[4], const float w25[4], float w[8]) {
  if (i >= 3 && i <= n - 4) {
    w[0] = w25[3]; w[1] = w75[3]; w[2] = w25[2]; w[3] = w75[2]; w[4] = w25[1]; w[5] = w75[1]; w[6] = w25[0]; w[7] = w75[0];
  } else {
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const int o = 2 * i - 3 + k;
      w[k] = (o >= 0 && o < 2 * n) ? cubic_up2_weight(o, i, n) : 0.f;
    }
  }
}
__global__ void __launch_bounds__(256) bicubic_up2_bwd_v4_kernel(const float* __restrict__ dy, float* __restrict__ dx, int B, int H, int W, int C) {
  const int Cv = C >> 2, Ho = 2 * H, Wo = 2 * W;
  const long long total = (long long)B * H * W * Cv;
  float w75[4], w25[4];
  cubic_coeffs(0.75f, w75); cubic_coeffs(0.25f, w25);
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(idx % Cv) << 2; long long r = idx / Cv;
    const int ix = (int)(r % W); r /= W; const int iy = (int)(r % H); const int b = (int)(r / H);
    float wy[8], wx[8];
    up2_bwd_weights(iy, H, w75, w25, wy);
    up2_bwd_weights(ix, W, w75, w25, wx);
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int oy = 2 * iy - 3 + i;
      if (oy < 0 || oy >= Ho) continue;
      const float* row = dy + ((size_t)b * Ho + oy) * Wo * C + c;
      float4 h = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const int ox = 2 * ix - 3 + j;
        if (ox < 0 || ox >= Wo) continue;
        const float4 q = *reinterpret_cast<const float4*>(row + (size_t)ox * C);
        h.x = fmaf(wx[j], q.x, h.x); h.y = fmaf(wx[j], q.y, h.y); h.z = fmaf(wx[j], q.z, h.z); h.w = fmaf(wx[j], q.w, h.w);
      }
      acc.x = fmaf(wy[i], h.x, acc.x); acc.y = fmaf(wy[i], h.y, acc.y); acc.z = fmaf(wy[i], h.z, acc.z); acc.w = fmaf(wy[i], h.w, acc.w);
    }
    *reinterpret_cast<float4*>(dx + (((size_t)b * H + iy) * W + ix) * C + c) = acc;
  }
}

// 2x2/2 max-pool backward: one thread per pooling window and 4 channels; the gradient goes to the first maximum in scan order
__global__ void __launch_bounds__(256) maxpool2_bwd_v4_kernel(const float* __restrict__ x, const float* __restrict__ dy, float* __restrict__ dx, int B, int H, int W, int C) {
  const int Cv = C >> 2, Ho = H / 2, Wo = W / 2, Hc = (H + 1) / 2, Wc = (W + 1) / 2;
  const long long total = (long long)B * Hc * Wc * Cv;
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(idx % Cv) << 2; long long r = idx / Cv;
    const int ox = (int)(r % Wc); r /= Wc; const int oy = (int)(r % Hc); const int b = (int)(r / Hc);
    const size_t p00 = (((size_t)b * H + 2 * oy) * W + 2 * ox) * C + c;
    const bool has_x1 = 2 * ox + 1 < W, has_y1 = 2 * oy + 1 < H;
    const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
    if (oy < Ho && ox < Wo) {
      const float4 v0 = *reinterpret_cast<const float4*>(x + p00), v1 = *reinterpret_cast<const float4*>(x + p00 + C);
      const float4 v2 = *reinterpret_cast<const float4*>(x + p00 + (size_t)W * C), v3 = *reinterpret_cast<const float4*>(x + p00 + (size_t)W * C + C);
      const float4 g = *reinterpret_cast<const float4*>(dy + (((size_t)b * Ho + oy) * Wo + ox) * C + c);
      float4 d0 = z, d1 = z, d2 = z, d3 = z;
#define GDN_POOL_ROUTE(f)                                                                       \
      { float m = v0.f; int a = 0; if (v1.f > m) { m = v1.f; a = 1; } if (v2.f > m) { m = v2.f; a = 2; } if (v3.f > m) { a = 3; } \
        d0.f = a == 0 ? g.f : 0.f; d1.f = a == 1 ? g.f : 0.f; d2.f = a == 2 ? g.f : 0.f; d3.f = a == 3 ? g.f : 0.f; }
      GDN_POOL_ROUTE(x) GDN_POOL_ROUTE(y) GDN_POOL_ROUTE(z) GDN_POOL_ROUTE(w)
#undef GDN_POOL_ROUTE
      *reinterpret_cast<float4*>(dx + p00) = d0; *reinterpret_cast<float4*>(dx + p00 + C) = d1;
      *reinterpret_cast<float4*>(dx + p00 + (size_t)W * C) = d2; *reinterpret_cast<float4*>(dx + p00 + (size_t)W * C + C) = d3;
    } else {   // odd H / W: the last row / column belongs to no window
      *reinterpret_cast<float4*>(dx + p00) = z;
      if (has_x1) *reinterpret_cast<float4*>(dx + p00 + C) = z;
      if (has_y1) *reinterpret_cast<float4*>(dx + p00 + (size_t)W * C) = z;
      if (has_x1 && has_y1) *reinterpret_cast<float4*>(dx + p00 + (size_t)W * C + C) = z;
    }
  }
}

// The same routing fused with the ReLU backward of the pooled map and the operand packing of the preceding convolution's data-gradient GEMM: x is the
// OUTPUT of conv + ReLU (VGG19 conv1_2 / conv2_2 / conv3_4, losses.py:58) and feeds only this pool, so dz = route(dy) * (x > 0) is needed only as a bf16
// [pixel][C] operand.  Replaces maxpool2_bwd (x read, fp32 dx written) + pack_actgrad (dx and x read again): 12 -> 6.5 bytes per element of x.
// One thread per pooling window and 8 channels; H, W even, C % 8 == 0.  Bit-identical to the two-pass path (same routing, same gate, same rounding).
__global__ void __launch_bounds__(256) maxpool2_bwd_relu_pack16_kernel(const float* __restrict__ x, const float* __restrict__ dy, __nv_bfloat16* __restrict__ dz,
                                                                       int B, int H, int W, int C) {
  const int Cv = C >> 3, Ho = H / 2, Wo = W / 2;
  const long long total = (long long)B * Ho * Wo * Cv;
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(idx % Cv) << 3; long long r = idx / Cv;
    const int ox = (int)(r % Wo); r /= Wo; const int oy = (int)(r % Ho); const int b = (int)(r / Ho);
    const size_t p00 = (((size_t)b * H + 2 * oy) * W + 2 * ox) * C + c;
    const size_t off[4] = {p00, p00 + C, p00 + (size_t)W * C, p00 + (size_t)W * C + C};
    float v[4][8], g[8];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const float4 a = __ldg(reinterpret_cast<const float4*>(x + off[k])), bq = __ldg(reinterpret_cast<const float4*>(x + off[k] + 4));
      v[k][0] = a.x; v[k][1] = a.y; v[k][2] = a.z; v[k][3] = a.w; v[k][4] = bq.x; v[k][5] = bq.y; v[k][6] = bq.z; v[k][7] = bq.w;
    }
    {
      const float* gp = dy + (((size_t)b * Ho + oy) * Wo + ox) * C + c;
      const float4 a = __ldg(reinterpret_cast<const float4*>(gp)), bq = __ldg(reinterpret_cast<const float4*>(gp + 4));
      g[0] = a.x; g[1] = a.y; g[2] = a.z; g[3] = a.w; g[4] = bq.x; g[5] = bq.y; g[6] = bq.z; g[7] = bq.w;
    }
    __align__(16) __nv_bfloat16 o[4][8];
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      float m = v[0][e]; int a = 0;
      if (v[1][e] > m) { m = v[1][e]; a = 1; }
      if (v[2][e] > m) { m = v[2][e]; a = 2; }
      if (v[3][e] > m) { m = v[3][e]; a = 3; }
      const float ge = m > 0.f ? g[e] : 0.f;                 // ReLU backward on the routed element (x is a ReLU output: x > 0 <=> pre-activation > 0)
#pragma unroll
      for (int k = 0; k < 4; ++k) o[k][e] = __float2bfloat16_rn(a == k ? ge : 0.f);
    }
#pragma unroll
    for (int k = 0; k < 4; ++k) *reinterpret_cast<uint4*>(dz + off[k]) = *reinterpret_cast<const uint4*>(o[k]);
  }
}

// ------------------------------------------------------------------ tap planes of a C -> 1 3x3 convolution (generator.py:228 after :225,:242-246)
// final(up2(u) + resize(s)) is linear and the resamplers act per channel, so the channel reduction of the 3x3 convolution can run BEFORE the
// resampling: z_t = sum_c w[c][t] u_c (a 1x1 convolution to 9 "tap planes" at low resolution), Z = up2(z) + resize(zs), and
//   y[p] = bias + sum_t Z_t[p + t]   (zero outside the grid: the convolution's zero padding)
// -- the 64-channel full-resolution tensor never exists (9 planes instead: 7x less traffic).  T = planes per pixel (9 used, padded to 12).
__global__ void __launch_bounds__(256) tap_shift_sum_kernel(const float* __restrict__ Z, int T, const float* __restrict__ bias, float* __restrict__ y, int B, int H, int W) {
  const long long total = (long long)B * H * W;
  const float b0 = bias ? __ldg(bias) : 0.f;
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
    const int j = (int)(idx % W); long long r = idx / W;
    const int i = (int)(r % H); const int b = (int)(r / H);
    float acc = b0;
#pragma unroll
    for (int a = 0; a < 3; ++a) {
      const int ii = i + a - 1;
      if (ii < 0 || ii >= H) continue;
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        const int jj = j + c - 1;
        if (jj < 0 || jj >= W) continue;
        acc += __ldg(Z + (((size_t)b * H + ii) * W + jj) * T + a * 3 + c);
      }
    }
    y[idx] = acc;
  }
}
// adjoint: dZ[b][i][j][a*3+c] = dy[b][i-a+1][j-c+1] (0 outside), planes 9..T-1 = 0
__global__ void __launch_bounds__(256) tap_shift_expand_kernel(const float* __restrict__ dy, float* __restrict__ dZ, int T, int B, int H, int W) {
  const long long total = (long long)B * H * W;
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
    const int j = (int)(idx % W); long long r = idx / W;
    const int i = (int)(r % H); const int b = (int)(r / H);
    float* o = dZ + (size_t)idx * T;
#pragma unroll
    for (int a = 0; a < 3; ++a) {
      const int ii = i - a + 1;
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        const int jj = j - c + 1;
        o[a * 3 + c] = (ii >= 0 && ii < H && jj >= 0 && jj < W) ? __ldg(dy + ((size_t)b * H + ii) * W + jj) : 0.f;
      }
    }
    for (int t = 9; t < T; ++t) o[t] = 0.f;
  }
}

// ------------------------------------------------------------------ bilinear resize to (Ho, Wo)
template <int VEC>
__global__ void __launch_bounds__(256) bilinear_fwd_kernel(const float* __restrict__ x, float* __restrict__ y, int B, int Hi, int Wi, int Ho, int Wo, int C,
                                                            float ry, float rx, int accumulate) {
  const int Cv = C / VEC;
  const long long total = (long long)B * Ho * Wo * Cv;
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
    int c = (int)(idx % Cv) * VEC; long long r = idx / Cv;
    int ox = (int)(r % Wo); r /= Wo; int oy = (int)(r % Ho); int b = (int)(r / Ho);
    int y0, y1, x0, x1; float ty, tx;
    linear_src(oy, ry, Hi, y0, y1, ty); linear_src(ox, rx, Wi, x0, x1, tx);
    const float w00 = (1.f - ty) * (1.f - tx), w01 = (1.f - ty) * tx, w10 = ty * (1.f - tx), w11 = ty * tx;
    const float* base = x + (size_t)b * Hi * Wi * C + c;
    const float* p00 = base + ((size_t)y0 * Wi + x0) * C; const float* p01 = base + ((size_t)y0 * Wi + x1) * C;
    const float* p10 = base + ((size_t)y1 * Wi + x0) * C; const float* p11 = base + ((size_t)y1 * Wi + x1) * C;
    float* o = y + (((size_t)b * Ho + oy) * Wo + ox) * C + c;
    if (VEC == 4) {
      float4 a = *reinterpret_cast<const float4*>(p00), bb = *reinterpret_cast<const float4*>(p01);
      float4 cc = *reinterpret_cast<const float4*>(p10), d = *reinterpret_cast<const float4*>(p11);
      float4 rr;
      rr.x = w00 * a.x + w01 * bb.x + w10 * cc.x + w11 * d.x; rr.y = w00 * a.y + w01 * bb.y + w10 * cc.y + w11 * d.y;
      rr.z = w00 * a.z + w01 * bb.z + w10 * cc.z + w11 * d.z; rr.w = w00 * a.w + w01 * bb.w + w10 * cc.w + w11 * d.w;
      if (accumulate) { float4 q = *reinterpret_cast<float4*>(o); rr.x += q.x; rr.y += q.y; rr.z += q.z; rr.w += q.w; }
      *reinterpret_cast<float4*>(o) = rr;
    } else {
      float rr = w00 * *p00 + w01 * *p01 + w10 * *p10 + w11 * *p11;
      if (accumulate) rr += *o;
      *o = rr;
    }
  }
}

__device__ __forceinline__ float linear_weight(int o, int i, float ratio, int n_in) {
  int i0, i1; float t;
  linear_src(o, ratio, n_in, i0, i1, t);
  float s = 0.f;
  if (i0 == i) s += 1.f - t;
  if (i1 == i) s += t;
  return s;
}

template <int VEC>
__global__ void __launch_bounds__(256) bilinear_bwd_kernel(const float* __restrict__ dy, float* __restrict__ dx, int B, int Hi, int Wi, int Ho, int Wo, int C,
                                                            float ry, float rx, int accumulate) {
  const int Cv = C / VEC;
  const long long total = (long long)B * Hi * Wi * Cv;
  const float iry = 1.f / ry, irx = 1.f / rx;
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
    int c = (int)(idx % Cv) * VEC; long long r = idx / Cv;
    int ix = (int)(r % Wi); r /= Wi; int iy = (int)(r % Hi); int b = (int)(r / Hi);
    // outputs whose source lies in (i-1, i+1): o in ((i-0.5)/ratio-0.5, (i+1.5)/ratio-0.5); widen by one and test exactly
    int oy_lo = (int)floorf((iy - 0.5f) * iry - 0.5f) - 1, oy_hi = (int)ceilf((iy + 1.5f) * iry - 0.5f) + 1;
    int ox_lo = (int)floorf((ix - 0.5f) * irx - 0.5f) - 1, ox_hi = (int)ceilf((ix + 1.5f) * irx - 0.5f) + 1;
    if (iy == 0) oy_lo = 0;          // the source clamp at 0 folds every earlier output onto row 0
    if (ix == 0) ox_lo = 0;
    oy_lo = oy_lo < 0 ? 0 : oy_lo; ox_lo = ox_lo < 0 ? 0 : ox_lo;
    oy_hi = oy_hi > Ho - 1 ? Ho - 1 : oy_hi; ox_hi = ox_hi > Wo - 1 ? Wo - 1 : ox_hi;
    if (iy == Hi - 1) oy_hi = Ho - 1;  // and the i1 clamp folds the tail onto the last row
    if (ix == Wi - 1) ox_hi = Wo - 1;
    float acc[VEC];
#pragma unroll
    for (int v = 0; v < VEC; ++v) acc[v] = 0.f;
    for (int oy = oy_lo; oy <= oy_hi; ++oy) {
      float wy = linear_weight(oy, iy, ry, Hi);
      if (wy == 0.f) continue;
      for (int ox = ox_lo; ox <= ox_hi; ++ox) {
        float wx = linear_weight(ox, ix, rx, Wi);
        if (wx == 0.f) continue;
        const float* p = dy + (((size_t)b * Ho + oy) * Wo + ox) * C + c;
        float wgt = wy * wx;
        if (VEC == 4) { float4 q = *reinterpret_cast<const float4*>(p); acc[0] = fmaf(wgt, q.x, acc[0]); acc[1] = fmaf(wgt, q.y, acc[1]); acc[2] = fmaf(wgt, q.z, acc[2]); acc[3] = fmaf(wgt, q.w, acc[3]); }
        else acc[0] = fmaf(wgt, *p, acc[0]);
      }
    }
    float* o = dx + (((size_t)b * Hi + iy) * Wi + ix) * C + c;
    if (VEC == 4) {
      float4 rr = make_float4(acc[0], acc[1], acc[2], acc[3]);
      if (accumulate) { float4 q = *reinterpret_cast<float4*>(o); rr.x += q.x; rr.y += q.y; rr.z += q.z; rr.w += q.w; }
      *reinterpret_cast<float4*>(o) = rr;
    } else {
      float rr = acc[0];
      if (accumulate) rr += *o;
      *o = rr;
    }
  }
}

// ------------------------------------------------------------------ bicubic 1/f down-sampling, NCHW -> NHWC slice
__device__ __forceinline__ float ld_as_float(const float* p) { return __ldg(p); }
__device__ __forceinline__ float ld_as_float(const __nv_bfloat16* p) { return __bfloat162float(*p); }
template <typename T>
__global__ void __launch_bounds__(256) bicubic_down_kernel(const T* __restrict__ x, float* __restrict__ y, int y_pitch, int B, int C, int Hi, int Wi,
                                                            int Ho, int Wo, float ratio) {
  const long long total = (long long)B * C * Ho * Wo;
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
    int ox = (int)(idx % Wo); long long r = idx / Wo;
    int oy = (int)(r % Ho); r /= Ho; int c = (int)(r % C); int b = (int)(r / C);
    int y0, x0; float ty, tx, wy[4], wx[4];
    cubic_src(oy, ratio, y0, ty); cubic_src(ox, ratio, x0, tx);
    cubic_coeffs(ty, wy); cubic_coeffs(tx, wx);
    const T* plane = x + ((size_t)b * C + c) * Hi * Wi;
    float acc = 0.f;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      int iy = clampi(y0 - 1 + i, 0, Hi - 1);
      float rowacc = 0.f;
#pragma unroll
      for (int j = 0; j < 4; ++j) rowacc = fmaf(wx[j], ld_as_float(plane + (size_t)iy * Wi + clampi(x0 - 1 + j, 0, Wi - 1)), rowacc);
      acc = fmaf(wy[i], rowacc, acc);
    }
    y[(((size_t)b * Ho + oy) * Wo + ox) * y_pitch + c] = acc;
  }
}

// ------------------------------------------------------------------ 2x2/2 max pooling
template <int VEC>
__global__ void __launch_bounds__(256) maxpool2_fwd_kernel(const float* __restrict__ x, float* __restrict__ y, int B, int H, int W, int C) {
  const int Cv = C / VEC, Ho = H / 2, Wo = W / 2;
  const long long total = (long long)B * Ho * Wo * Cv;
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
    int c = (int)(idx % Cv) * VEC; long long r = idx / Cv;
    int ox = (int)(r % Wo); r /= Wo; int oy = (int)(r % Ho); int b = (int)(r / Ho);
    const float* p = x + (((size_t)b * H + 2 * oy) * W + 2 * ox) * C + c;
    float* o = y + (((size_t)b * Ho + oy) * Wo + ox) * C + c;
    if (VEC == 4) {
      float4 a = *reinterpret_cast<const float4*>(p), bb = *reinterpret_cast<const float4*>(p + C);
      float4 cc = *reinterpret_cast<const float4*>(p + (size_t)W * C), d = *reinterpret_cast<const float4*>(p + (size_t)W * C + C);
      *reinterpret_cast<float4*>(o) = make_float4(fmaxf(fmaxf(a.x, bb.x), fmaxf(cc.x, d.x)), fmaxf(fmaxf(a.y, bb.y), fmaxf(cc.y, d.y)),
                                                  fmaxf(fmaxf(a.z, bb.z), fmaxf(cc.z, d.z)), fmaxf(fmaxf(a.w, bb.w), fmaxf(cc.w, d.w)));
    } else {
      *o = fmaxf(fmaxf(p[0], p[C]), fmaxf(p[(size_t)W * C], p[(size_t)W * C + C]));
    }
  }
}
// Same gradient for a vertical strip of T input rows per thread: the horizontally filtered dy rows are shared by the T outputs of the strip
// (2T + 6 rows of 8 loads instead of T x 64 loads: 22 instead of 64 128-bit loads per output at T = 8; the one-pixel kernel is bound by
// L1 requests, 1.2 TB/s).  Strips that touch the top / bottom border (clamped taps fold together there) take the generic per-pixel path.
template <int T>
__global__ void __launch_bounds__(256) bicubic_up2_bwd_strip_kernel(const float* __restrict__ dy, float* __restrict__ dx, int B, int H, int W, int C) {
  const int Cv = C >> 2, Ho = 2 * H, Wo = 2 * W, strips = (H + T - 1) / T;
  const long long total = (long long)B * strips * W * Cv;
  float w75[4], w25[4];
  cubic_coeffs(0.75f, w75); cubic_coeffs(0.25f, w25);
  const float wyc[8] = {w25[3], w75[3], w25[2], w75[2], w25[1], w75[1], w25[0], w75[0]};     // interior rows (see up2_bwd_weights)
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(idx % Cv) << 2; long long r = idx / Cv;
    const int ix = (int)(r % W); r /= W; const int st = (int)(r % strips); const int b = (int)(r / strips);
    const int iy0 = st * T;
    float wx[8];
    up2_bwd_weights(ix, W, w75, w25, wx);
    const float* base = dy + (size_t)b * Ho * Wo * C + c;
    if (iy0 >= 3 && iy0 + T - 1 <= H - 4) {
      float4 acc[T];
#pragma unroll
      for (int k = 0; k < T; ++k) acc[k] = make_float4(0.f, 0.f, 0.f, 0.f);
      // rows are software-pipelined two deep: the kernel is bound by DRAM latency x the 2T+6 sequential rows of a strip (ncu: 24 % warps
      // active, every FMA waiting on its loads), so the loads of the next two rows are in flight while a row is filtered
      const float* row0 = base + (size_t)(2 * iy0 - 3) * Wo * C;
      const size_t rstride = (size_t)Wo * C;
      bool okx[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) { const int ox = 2 * ix - 3 + j; okx[j] = ox >= 0 && ox < Wo; }
      const float* col0 = row0 + (ptrdiff_t)(2 * ix - 3) * C;
      float4 q[3][8];
#pragma unroll
      for (int pre = 0; pre < 2; ++pre)
#pragma unroll
        for (int j = 0; j < 8; ++j) q[pre][j] = okx[j] ? __ldg(reinterpret_cast<const float4*>(col0 + pre * rstride + (size_t)j * C)) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
      for (int rr = 0; rr < 2 * T + 6; ++rr) {
        if (rr + 2 < 2 * T + 6) {
#pragma unroll
          for (int j = 0; j < 8; ++j)
            q[(rr + 2) % 3][j] = okx[j] ? __ldg(reinterpret_cast<const float4*>(col0 + (rr + 2) * rstride + (size_t)j * C)) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
        float4 h = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float4 v = q[rr % 3][j];
          h.x = fmaf(wx[j], v.x, h.x); h.y = fmaf(wx[j], v.y, h.y); h.z = fmaf(wx[j], v.z, h.z); h.w = fmaf(wx[j], v.w, h.w);
        }
#pragma unroll
        for (int k = 0; k < T; ++k) {
          const int i = rr - 2 * k;                      // compile-time after unrolling
          if (i >= 0 && i < 8) { acc[k].x = fmaf(wyc[i], h.x, acc[k].x); acc[k].y = fmaf(wyc[i], h.y, acc[k].y); acc[k].z = fmaf(wyc[i], h.z, acc[k].z); acc[k].w = fmaf(wyc[i], h.w, acc[k].w); }
        }
      }
#pragma unroll
      for (int k = 0; k < T; ++k) *reinterpret_cast<float4*>(dx + (((size_t)b * H + iy0 + k) * W + ix) * C + c) = acc[k];
    } else {
      for (int k = 0; k < T && iy0 + k < H; ++k) {
        const int iy = iy0 + k;
        float wy[8];
        up2_bwd_weights(iy, H, w75, w25, wy);
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const int oy = 2 * iy - 3 + i;
          if (oy < 0 || oy >= Ho) continue;
          const float* row = base + (size_t)oy * Wo * C;
          float4 h = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const int ox = 2 * ix - 3 + j;
            if (ox < 0 || ox >= Wo) continue;
            const float4 q = __ldg(reinterpret_cast<const float4*>(row + (size_t)ox * C));
            h.x = fmaf(wx[j], q.x, h.x); h.y = fmaf(wx[j], q.y, h.y); h.z = fmaf(wx[j], q.z, h.z); h.w = fmaf(wx[j], q.w, h.w);
          }
          acc.x = fmaf(wy[i], h.x, acc.x); acc.y = fmaf(wy[i], h.y, acc.y); acc.z = fmaf(wy[i], h.z, acc.z); acc.w = fmaf(wy[i], h.w, acc.w);
        }
        *reinterpret_cast<float4*>(dx + (((size_t)b * H + iy) * W + ix) * C + c) = acc;
      }
    }
  }
}

// bf16 activations (the frozen VGG19 branch keeps its untapped feature maps in bf16 only): 8 channels = 16 bytes per thread
__device__ __forceinline__ void bf16x8_to_float(const uint4& q, float* f) {
  const __nv_bfloat16* h = reinterpret_cast<const __nv_bfloat16*>(&q);
#pragma unroll
  for (int e = 0; e < 8; ++e) f[e] = __bfloat162float(h[e]);
}
__global__ void __launch_bounds__(256) maxpool2_fwd_bf16_kernel(const __nv_bfloat16* __restrict__ x, __nv_bfloat16* __restrict__ y, int B, int H, int W, int C) {
  const int Cv = C >> 3, Ho = H / 2, Wo = W / 2;
  const long long total = (long long)B * Ho * Wo * Cv;
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(idx % Cv) << 3; long long r = idx / Cv;
    const int ox = (int)(r % Wo); r /= Wo; const int oy = (int)(r % Ho); const int b = (int)(r / Ho);
    const __nv_bfloat16* p = x + (((size_t)b * H + 2 * oy) * W + 2 * ox) * C + c;
    float a[8], bb[8], cc[8], d[8];
    bf16x8_to_float(__ldg(reinterpret_cast<const uint4*>(p)), a); bf16x8_to_float(__ldg(reinterpret_cast<const uint4*>(p + C)), bb);
    bf16x8_to_float(__ldg(reinterpret_cast<const uint4*>(p + (size_t)W * C)), cc); bf16x8_to_float(__ldg(reinterpret_cast<const uint4*>(p + (size_t)W * C + C)), d);
    __align__(16) __nv_bfloat16 o[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) o[e] = __float2bfloat16_rn(fmaxf(fmaxf(a[e], bb[e]), fmaxf(cc[e], d[e])));
    *reinterpret_cast<uint4*>(y + (((size_t)b * Ho + oy) * Wo + ox) * C + c) = *reinterpret_cast<const uint4*>(o);
  }
}
// dx (fp32) of the bf16 pooling input: the gradient goes to the first maximum in scan order; even H, W
__global__ void __launch_bounds__(256) maxpool2_bwd_bf16_kernel(const __nv_bfloat16* __restrict__ x, const float* __restrict__ dy, float* __restrict__ dx, int B, int H, int W, int C) {
  const int Cv = C >> 2, Ho = H / 2, Wo = W / 2;
  const long long total = (long long)B * Ho * Wo * Cv;
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(idx % Cv) << 2; long long r = idx / Cv;
    const int ox = (int)(r % Wo); r /= Wo; const int oy = (int)(r % Ho); const int b = (int)(r / Ho);
    const size_t p00 = (((size_t)b * H + 2 * oy) * W + 2 * ox) * C + c;
    float v[4][4];
    const size_t offs[4] = {p00, p00 + (size_t)C, p00 + (size_t)W * C, p00 + (size_t)W * C + C};
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const uint2 q = __ldg(reinterpret_cast<const uint2*>(x + offs[k]));
      const __nv_bfloat16* h = reinterpret_cast<const __nv_bfloat16*>(&q);
#pragma unroll
      for (int e = 0; e < 4; ++e) v[k][e] = __bfloat162float(h[e]);
    }
    const float4 g = *reinterpret_cast<const float4*>(dy + (((size_t)b * Ho + oy) * Wo + ox) * C + c);
    const float gv[4] = {g.x, g.y, g.z, g.w};
    float d[4][4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      float m = v[0][e]; int a = 0;
      if (v[1][e] > m) { m = v[1][e]; a = 1; }
      if (v[2][e] > m) { m = v[2][e]; a = 2; }
      if (v[3][e] > m) { a = 3; }
#pragma unroll
      for (int k = 0; k < 4; ++k) d[k][e] = a == k ? gv[e] : 0.f;
    }
#pragma unroll
    for (int k = 0; k < 4; ++k) *reinterpret_cast<float4*>(dx + offs[k]) = make_float4(d[k][0], d[k][1], d[k][2], d[k][3]);
  }
}
// gradient goes to the first maximum in scan order (ATen max_pool2d backward)
__global__ void __launch_bounds__(256) maxpool2_bwd_kernel(const float* __restrict__ x, const float* __restrict__ dy, float* __restrict__ dx, int B, int H, int W, int C) {
  const int Ho = H / 2, Wo = W / 2;
  const long long total = (long long)B * H * W * C;
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
    int c = (int)(idx % C); long long r = idx / C;
    int ix = (int)(r % W); r /= W; int iy = (int)(r % H); int b = (int)(r / H);
    int oy = iy >> 1, ox = ix >> 1;
    float g = 0.f;
    if (oy < Ho && ox < Wo) {
      const float* p = x + (((size_t)b * H + 2 * oy) * W + 2 * ox) * C + c;
      float v[4] = {p[0], p[C], p[(size_t)W * C], p[(size_t)W * C + C]};
      int arg = 0; float m = v[0];
#pragma unroll
      for (int k = 1; k < 4; ++k) if (v[k] > m) { m = v[k]; arg = k; }
      if (arg == ((iy & 1) * 2 + (ix & 1))) g = dy[(((size_t)b * Ho + oy) * Wo + ox) * C + c];
    }
    dx[idx] = g;
  }
}

static inline int grid_for(long long total) {
  long long b = cdiv(total, 256);
  return (int)(b < 16 * kNumSMs ? (b > 0 ? b : 1) : 16 * kNumSMs);
}
static inline bool al16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }
}  // namespace gdn

using namespace gdn;

extern "C" int gdn_bicubic_up2_fwd(const float* x, float* y, int B, int H, int W, int C, gdn_stream_t s) {
  GDN_CHECK_ARG(x && y && B > 0 && H > 0 && W > 0 && C > 0);
  if (C % 4 == 0 && al16(x) && al16(y)) bicubic_up2_fwd_v4_kernel<0><<<grid_for((long long)B * H * W * (C / 4)), 256, 0, as_stream(s)>>>(x, y, B, H, W, C, nullptr, 0, 0);
  else bicubic_up2_fwd_kernel<1><<<grid_for((long long)B * 4 * H * W * C), 256, 0, as_stream(s)>>>(x, y, B, H, W, C);
  GDN_CHECK_LAUNCH();
  return GDN_OK;
}
extern "C" int gdn_bicubic_up2_bilinear_add_fwd(const float* x, const float* skip, float* y, int B, int H, int W, int Hs, int Ws, int C, gdn_stream_t s) {
  GDN_CHECK_ARG(x && skip && y && B > 0 && H > 0 && W > 0 && Hs > 0 && Ws > 0 && C > 0 && C % 4 == 0 && al16(x) && al16(y) && al16(skip));
  bicubic_up2_fwd_v4_kernel<1><<<grid_for((long long)B * H * W * (C / 4)), 256, 0, as_stream(s)>>>(x, y, B, H, W, C, skip, Hs, Ws);
  GDN_CHECK_LAUNCH();
  return GDN_OK;
}
extern "C" int gdn_tap_shift_sum(const float* Z, int T, const float* bias, float* y, int B, int H, int W, gdn_stream_t s) {
  GDN_CHECK_ARG(Z && y && T >= 9 && B > 0 && H > 0 && W > 0);
  tap_shift_sum_kernel<<<grid_for((long long)B * H * W), 256, 0, as_stream(s)>>>(Z, T, bias, y, B, H, W);
  GDN_CHECK_LAUNCH();
  return GDN_OK;
}
extern "C" int gdn_tap_shift_expand(const float* dy, float* dZ, int T, int B, int H, int W, gdn_stream_t s) {
  GDN_CHECK_ARG(dy && dZ && T >= 9 && B > 0 && H > 0 && W > 0);
  tap_shift_expand_kernel<<<grid_for((long long)B * H * W), 256, 0, as_stream(s)>>>(dy, dZ, T, B, H, W);
  GDN_CHECK_LAUNCH();
  return GDN_OK;
}
extern "C" int gdn_bicubic_up2_bwd(const float* dy, float* dx, int B, int H, int W, int C, gdn_stream_t s) {
  GDN_CHECK_ARG(dy && dx && B > 0 && H > 0 && W > 0 && C > 0);
  if (C % 4 == 0 && al16(dy) && al16(dx) && H >= 24)
    bicubic_up2_bwd_strip_kernel<8><<<grid_for((long long)B * ((H + 7) / 8) * W * (C / 4)), 256, 0, as_stream(s)>>>(dy, dx, B, H, W, C);
  else if (C % 4 == 0 && al16(dy) && al16(dx)) bicubic_up2_bwd_v4_kernel<<<grid_for((long long)B * H * W * (C / 4)), 256, 0, as_stream(s)>>>(dy, dx, B, H, W, C);
  else bicubic_up2_bwd_kernel<1><<<grid_for((long long)B * H * W * C), 256, 0, as_stream(s)>>>(dy, dx, B, H, W, C);
  GDN_CHECK_LAUNCH();
  return GDN_OK;
}
extern "C" int gdn_bilinear_fwd(const float* x, float* y, int B, int Hi, int Wi, int Ho, int Wo, int C, int accumulate, gdn_stream_t s) {
  GDN_CHECK_ARG(x && y && B > 0 && Hi > 0 && Wi > 0 && Ho > 0 && Wo > 0 && C > 0);
  float ry = (float)Hi / (float)Ho, rx = (float)Wi / (float)Wo;
  if (C % 4 == 0 && al16(x) && al16(y)) bilinear_fwd_kernel<4><<<grid_for((long long)B * Ho * Wo * (C / 4)), 256, 0, as_stream(s)>>>(x, y, B, Hi, Wi, Ho, Wo, C, ry, rx, accumulate);
  else bilinear_fwd_kernel<1><<<grid_for((long long)B * Ho * Wo * C), 256, 0, as_stream(s)>>>(x, y, B, Hi, Wi, Ho, Wo, C, ry, rx, accumulate);
  GDN_CHECK_LAUNCH();
  return GDN_OK;
}
extern "C" int gdn_bilinear_bwd(const float* dy, float* dx, int B, int Hi, int Wi, int Ho, int Wo, int C, int accumulate, gdn_stream_t s) {
  GDN_CHECK_ARG(dy && dx && B > 0 && Hi > 0 && Wi > 0 && Ho > 0 && Wo > 0 && C > 0);
  float ry = (float)Hi / (float)Ho, rx = (float)Wi / (float)Wo;
  if (C % 4 == 0 && al16(dy) && al16(dx)) bilinear_bwd_kernel<4><<<grid_for((long long)B * Hi * Wi * (C / 4)), 256, 0, as_stream(s)>>>(dy, dx, B, Hi, Wi, Ho, Wo, C, ry, rx, accumulate);
  else bilinear_bwd_kernel<1><<<grid_for((long long)B * Hi * Wi * C), 256, 0, as_stream(s)>>>(dy, dx, B, Hi, Wi, Ho, Wo, C, ry, rx, accumulate);
  GDN_CHECK_LAUNCH();
  return GDN_OK;
}
extern "C" int gdn_bicubic_down_nchw_to_nhwc(const float* x, float* y, int y_pitch, int y_c0, int B, int C, int Hi, int Wi, int f, gdn_stream_t s) {
  GDN_CHECK_ARG(x && y && B > 0 && C > 0 && Hi > 0 && Wi > 0 && (f == 2 || f == 4) && y_pitch >= y_c0 + C);
  int Ho = Hi / f, Wo = Wi / f;   // floor(in * 1/f)
  GDN_CHECK_ARG(Ho > 0 && Wo > 0);
  bicubic_down_kernel<float><<<grid_for((long long)B * C * Ho * Wo), 256, 0, as_stream(s)>>>(x, y + y_c0, y_pitch, B, C, Hi, Wi, Ho, Wo, (float)f);
  GDN_CHECK_LAUNCH();
  return GDN_OK;
}
extern "C" int gdn_bicubic_down_nchw_to_nhwc_bf16(const uint16_t* x, float* y, int y_pitch, int y_c0, int B, int C, int Hi, int Wi, int f, gdn_stream_t s) {
  GDN_CHECK_ARG(x && y && B > 0 && C > 0 && Hi > 0 && Wi > 0 && (f == 2 || f == 4) && y_pitch >= y_c0 + C);
  int Ho = Hi / f, Wo = Wi / f;
  GDN_CHECK_ARG(Ho > 0 && Wo > 0);
  bicubic_down_kernel<__nv_bfloat16><<<grid_for((long long)B * C * Ho * Wo), 256, 0, as_stream(s)>>>(reinterpret_cast<const __nv_bfloat16*>(x), y + y_c0, y_pitch, B, C, Hi, Wi,
                                                                                                    Ho, Wo, (float)f);
  GDN_CHECK_LAUNCH();
  return GDN_OK;
}
extern "C" int gdn_maxpool2_fwd(const float* x, float* y, int B, int H, int W, int C, gdn_stream_t s) {
  GDN_CHECK_ARG(x && y && B > 0 && H >= 2 && W >= 2 && C > 0);
  if (C % 4 == 0 && al16(x) && al16(y)) maxpool2_fwd_kernel<4><<<grid_for((long long)B * (H / 2) * (W / 2) * (C / 4)), 256, 0, as_stream(s)>>>(x, y, B, H, W, C);
  else maxpool2_fwd_kernel<1><<<grid_for((long long)B * (H / 2) * (W / 2) * C), 256, 0, as_stream(s)>>>(x, y, B, H, W, C);
  GDN_CHECK_LAUNCH();
  return GDN_OK;
}
extern "C" int gdn_maxpool2_fwd_bf16(const uint16_t* x, uint16_t* y, int B, int H, int W, int C, gdn_stream_t s) {
  GDN_CHECK_ARG(x && y && B > 0 && H >= 2 && W >= 2 && C > 0 && C % 8 == 0 && al16(x) && al16(y));
  maxpool2_fwd_bf16_kernel<<<grid_for((long long)B * (H / 2) * (W / 2) * (C / 8)), 256, 0, as_stream(s)>>>(reinterpret_cast<const __nv_bfloat16*>(x),
                                                                                                     reinterpret_cast<__nv_bfloat16*>(y), B, H, W, C);
  GDN_CHECK_LAUNCH();
  return GDN_OK;
}
extern "C" int gdn_maxpool2_bwd_bf16(const uint16_t* x, const float* dy, float* dx, int B, int H, int W, int C, gdn_stream_t s) {
  GDN_CHECK_ARG(x && dy && dx && B > 0 && H >= 2 && W >= 2 && H % 2 == 0 && W % 2 == 0 && C > 0 && C % 8 == 0 && al16(x) && al16(dy) && al16(dx));
  maxpool2_bwd_bf16_kernel<<<grid_for((long long)B * (H / 2) * (W / 2) * (C / 4)), 256, 0, as_stream(s)>>>(reinterpret_cast<const __nv_bfloat16*>(x), dy, dx, B, H, W, C);
  GDN_CHECK_LAUNCH();
  return GDN_OK;
}
extern "C" int gdn_maxpool2_bwd_relu_pack16(const float* x, const float* dy, uint16_t* dz16, int B, int H, int W, int C, gdn_stream_t s) {
  GDN_CHECK_ARG(x && dy && dz16 && B > 0 && H >= 2 && W >= 2 && H % 2 == 0 && W % 2 == 0 && C > 0 && C % 8 == 0 && al16(x) && al16(dy) && al16(dz16));
  maxpool2_bwd_relu_pack16_kernel<<<grid_for((long long)B * (H / 2) * (W / 2) * (C / 8)), 256, 0, as_stream(s)>>>(x, dy, reinterpret_cast<__nv_bfloat16*>(dz16), B, H, W, C);
  GDN_CHECK_LAUNCH();
  return GDN_OK;
}
extern "C" int gdn_maxpool2_bwd(const float* x, const float* dy, float* dx, int B, int H, int W, int C, gdn_stream_t s) {
  GDN_CHECK_ARG(x && dy && dx && B > 0 && H >= 2 && W >= 2 && C > 0);
  if (C % 4 == 0 && al16(x) && al16(dy) && al16(dx))
    maxpool2_bwd_v4_kernel<<<grid_for((long long)B * ((H + 1) / 2) * ((W + 1) / 2) * (C / 4)), 256, 0, as_stream(s)>>>(x, dy, dx, B, H, W, C);
  else
    maxpool2_bwd_kernel<<<grid_for((long long)B * H * W * C), 256, 0, as_stream(s)>>>(x, dy, dx, B, H, W, C);
  GDN_CHECK_LAUNCH();
  return GDN_OK;
}
