// "Thin" 3x3 convolutions: one side has a single channel, the other C channels (C = 32, 64 or 128).  They are HBM-bound
// (D conv1 1->64, VGG19 conv1_1 on the channel-summed weight, the generator's final 64->1 conv, and their
// gradients), so they run as coalesced CUDA-core kernels that read or write the wide tensor exactly once instead of as padded
// tensor-core tiles.  Call sites in the reference: models/discriminator.py:62 (conv1), models/generator.py:228 (final),
// models/losses.py:58,64-65 (conv1_1 on x.repeat(1,3,1,1)).  All kernels share one geometry: a "wide" NHWC tensor V [B,Hv,Wv,C]
// and a single-channel field S [B,Hs,Ws], related by  s-coordinate = v-coordinate * stride + (k - pad)  for tap k.
//   expand : V[p][c]  = act(sum_k S[p*stride + k - pad] * w[c][k] + bias[c]) (+ res)         1 -> C forward; data gradient of C -> 1
//   reduce : S[q]     = sum_k sum_c V[p(q,k)][c] * w[c][k] (+ bias) (+ res)                  C -> 1 forward; data gradient of 1 -> C
//   wgrad  : dw[c][k] = sum_p V[p][c] * S[p*stride + k - pad]                                weight gradient of both
// Design rules: a thread's channel group is fixed, so its filter taps live in registers; each thread handles 4 x-adjacent pixels per
// iteration (four independent 16-byte transactions in flight, S window shared); all index arithmetic is 32-bit.
// A flipped tap order (the "transposed" uses) is folded into the register copy of the weights:  tap k at offset pad - k  ==  tap
// 8 - k at offset k - (2 - pad).
#include <cuda_bf16.h>
#include "common.cuh"

namespace gdn {
namespace thin {

struct Geo { int B, Hv, Wv, C, Hs, Ws, stride, pad; };

// ---------------------------------------------------------------------------------------------------------------- expand
// STRIDE == 0: run-time stride, one pixel per iteration.  C/4 lanes per pixel, one float4 of channels each.
template <int STRIDE, int PX>
__global__ void __launch_bounds__(256) expand_kernel(const float* __restrict__ S, const float* __restrict__ w, const float* __restrict__ bias, float* V, int v_pitch,
                                                     const float* res, int res_pitch,   /* res may alias V */ Geo g, int flip, int act, float slope,
                                                     __nv_bfloat16* __restrict__ V16 /* optional bf16 copy [pixel][C] */) {
  constexpr int NC = STRIDE ? (PX - 1) * STRIDE + 3 : 3;
  const int lanes = g.C >> 2;
  const int c = (threadIdx.x % lanes) << 2;
  const int stride = STRIDE ? STRIDE : g.stride;
  const int pad = flip ? 2 - g.pad : g.pad;
  float wr[4][9];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int k = 0; k < 9; ++k) wr[i][k] = __ldg(w + (c + i) * 9 + (flip ? 8 - k : k));
  const float4 bv = bias ? __ldg(reinterpret_cast<const float4*>(bias + c)) : make_float4(0.f, 0.f, 0.f, 0.f);
  const int gx = g.Wv / PX;
  const int total = g.B * g.Hv * gx;
  const int gpp = 256 / lanes;                                  // pixel groups per block pass
  // the kernel was instruction-issue bound (ncu: 75 % issue slots, 3.3 TB/s): packed fp32x2 FMAs (sm_100 FFMA2), an interior fast path
  // without per-element bounds checks and 32-bit offsets take it to the write bandwidth
  float2 w01[9], w23[9];
#pragma unroll
  for (int k = 0; k < 9; ++k) { w01[k] = make_float2(wr[0][k], wr[1][k]); w23[k] = make_float2(wr[2][k], wr[3][k]); }
  for (int gi = blockIdx.x * gpp + threadIdx.x / lanes; gi < total; gi += gridDim.x * gpp) {
    const int xg = gi % gx, r = gi / gx, y = r % g.Hv, b = r / g.Hv;
    const int x0 = xg * PX;
    const float* sb = S + (size_t)b * g.Hs * g.Ws;
    const int sx0 = x0 * stride - pad, sy0 = y * stride - pad;
    float sw[3][NC];
    if (sy0 >= 0 && sy0 + 2 < g.Hs && sx0 >= 0 && sx0 + NC <= g.Ws) {          // interior: the whole window is inside the field
      const float* s0 = sb + (size_t)sy0 * g.Ws + sx0;
#pragma unroll
      for (int kh = 0; kh < 3; ++kh)
#pragma unroll
        for (int j = 0; j < NC; ++j) sw[kh][j] = __ldg(s0 + kh * g.Ws + j);
    } else {
#pragma unroll
      for (int kh = 0; kh < 3; ++kh) {
        const int sy = sy0 + kh;
        const bool rowok = sy >= 0 && sy < g.Hs;
        const float* sr = sb + (size_t)(rowok ? sy : 0) * g.Ws;
#pragma unroll
        for (int j = 0; j < NC; ++j) {
          const int sx = sx0 + j;
          sw[kh][j] = (rowok && sx >= 0 && sx < g.Ws) ? __ldg(sr + sx) : 0.f;
        }
      }
    }
    const size_t p = ((size_t)b * g.Hv + y) * g.Wv + x0;
#pragma unroll
    for (int i = 0; i < PX; ++i) {
      float2 a01 = make_float2(bv.x, bv.y), a23 = make_float2(bv.z, bv.w);
#pragma unroll
      for (int kh = 0; kh < 3; ++kh)
#pragma unroll
        for (int kw = 0; kw < 3; ++kw) {
          const float sv = sw[kh][(STRIDE ? i * STRIDE : 0) + kw];
          const float2 s2 = make_float2(sv, sv);
          a01 = __ffma2_rn(s2, w01[kh * 3 + kw], a01);
          a23 = __ffma2_rn(s2, w23[kh * 3 + kw], a23);
        }
      float4 acc = make_float4(a01.x, a01.y, a23.x, a23.y);
      if (act == GDN_ACT_RELU) { acc.x = fmaxf(acc.x, 0.f); acc.y = fmaxf(acc.y, 0.f); acc.z = fmaxf(acc.z, 0.f); acc.w = fmaxf(acc.w, 0.f); }
      else if (act == GDN_ACT_LRELU) { acc.x = fmaxf(acc.x, acc.x * slope); acc.y = fmaxf(acc.y, acc.y * slope); acc.z = fmaxf(acc.z, acc.z * slope); acc.w = fmaxf(acc.w, acc.w * slope); }
      if (res) { const float4 rr = *reinterpret_cast<const float4*>(res + (p + i) * res_pitch + c); acc.x += rr.x; acc.y += rr.y; acc.z += rr.z; acc.w += rr.w; }
      if (V) *reinterpret_cast<float4*>(V + (p + i) * v_pitch + c) = acc;
      if (V16) {
        const __nv_bfloat162 h0 = __floats2bfloat162_rn(acc.x, acc.y), h1 = __floats2bfloat162_rn(acc.z, acc.w);
        *reinterpret_cast<uint2*>(V16 + (p + i) * g.C + c) = make_uint2(*reinterpret_cast<const uint32_t*>(&h0), *reinterpret_cast<const uint32_t*>(&h1));
      }
    }
  }
}


// ---------------------------------------------------------------------------------------------------------------- tapped 1 -> C conv + ReLU of an image PAIR
// The perceptual loss taps relu1_1 (losses.py:60-72 with feature_layers containing 1) and conv1_1 has ONE input channel here (the channel-summed
// weight, losses.py:64-65): both 64-channel maps are 9 FMAs per element away from the two single-channel images.  One pass over the two images
//   * writes both maps ONLY as the bf16 operands of conv1_2 (V16a generated, V16b target),
//   * accumulates the L1 term  sum | relu(conv Sa) - relu(conv Sb) |,
//   * and keeps what the backward pass needs of the two fp32 maps -- 2 bits per element: 0 = ReLU closed (fa <= 0), else 2 + sign(fa' - fb') --
//     as one byte per (pixel, 4 channels): 67 MB instead of 2 x 1.07 GB at 256x512, batch 32.
// The fp32 maps (2 x 1.07 GB written), the L1 pass over them (2.1 GB read, 1 GB of gradient written and re-read) never exist.  Same FMA order as
// expand_kernel => gate, sign and bf16 operands are bitwise those of the stored-map path.
template <int PX>
__global__ void __launch_bounds__(256, 2) tap_pair_kernel(const float* __restrict__ Sa, const float* __restrict__ Sb, const float* __restrict__ w,
                                                          const float* __restrict__ bias, Geo g, double* __restrict__ partial,
                                                          __nv_bfloat16* __restrict__ V16a, __nv_bfloat16* __restrict__ V16b, uint8_t* __restrict__ mask) {
  constexpr int NC = PX + 2;
  const int lanes = g.C >> 2;
  const int lc = threadIdx.x % lanes;
  const int c = lc << 2;
  float2 w01[9], w23[9];
#pragma unroll
  for (int k = 0; k < 9; ++k) {
    w01[k] = make_float2(__ldg(w + (c + 0) * 9 + k), __ldg(w + (c + 1) * 9 + k));
    w23[k] = make_float2(__ldg(w + (c + 2) * 9 + k), __ldg(w + (c + 3) * 9 + k));
  }
  const float4 bv = bias ? __ldg(reinterpret_cast<const float4*>(bias + c)) : make_float4(0.f, 0.f, 0.f, 0.f);
  const int gx = g.Wv / PX;
  const int total = g.B * g.Hv * gx;
  const int gpp = 256 / lanes;
  float acc = 0.f; double dacc = 0.0; int cnt = 0;
  for (int gi = blockIdx.x * gpp + threadIdx.x / lanes; gi < total; gi += gridDim.x * gpp) {
    const int xg = gi % gx, r = gi / gx, y = r % g.Hv, b = r / g.Hv;
    const int x0 = xg * PX;
    const size_t fo = (size_t)b * g.Hs * g.Ws;
    const int sx0 = x0 - 1, sy0 = y - 1;
    float sa[3][NC], sb[3][NC];
    if (sy0 >= 0 && sy0 + 2 < g.Hs && sx0 >= 0 && sx0 + NC <= g.Ws) {
      const size_t o0 = fo + (size_t)sy0 * g.Ws + sx0;
#pragma unroll
      for (int kh = 0; kh < 3; ++kh)
#pragma unroll
        for (int j = 0; j < NC; ++j) { sa[kh][j] = __ldg(Sa + o0 + kh * g.Ws + j); sb[kh][j] = __ldg(Sb + o0 + kh * g.Ws + j); }
    } else {
#pragma unroll
      for (int kh = 0; kh < 3; ++kh) {
        const int sy = sy0 + kh;
        const bool rowok = sy >= 0 && sy < g.Hs;
        const size_t ro = fo + (size_t)(rowok ? sy : 0) * g.Ws;
#pragma unroll
        for (int j = 0; j < NC; ++j) {
          const int sx = sx0 + j;
          const bool ok = rowok && sx >= 0 && sx < g.Ws;
          sa[kh][j] = ok ? __ldg(Sa + ro + sx) : 0.f; sb[kh][j] = ok ? __ldg(Sb + ro + sx) : 0.f;
        }
      }
    }
    const size_t p = ((size_t)b * g.Hv + y) * g.Wv + x0;
#pragma unroll
    for (int i = 0; i < PX; ++i) {
      float2 a01 = make_float2(bv.x, bv.y), a23 = make_float2(bv.z, bv.w), b01 = a01, b23 = a23;
#pragma unroll
      for (int kh = 0; kh < 3; ++kh)
#pragma unroll
        for (int kw = 0; kw < 3; ++kw) {
          const float2 s2 = make_float2(sa[kh][i + kw], sa[kh][i + kw]), t2 = make_float2(sb[kh][i + kw], sb[kh][i + kw]);
          a01 = __ffma2_rn(s2, w01[kh * 3 + kw], a01); a23 = __ffma2_rn(s2, w23[kh * 3 + kw], a23);
          b01 = __ffma2_rn(t2, w01[kh * 3 + kw], b01); b23 = __ffma2_rn(t2, w23[kh * 3 + kw], b23);
        }
      const float fa[4] = {a01.x, a01.y, a23.x, a23.y};
      const float ra[4] = {fmaxf(a01.x, 0.f), fmaxf(a01.y, 0.f), fmaxf(a23.x, 0.f), fmaxf(a23.y, 0.f)};
      const float rb[4] = {fmaxf(b01.x, 0.f), fmaxf(b01.y, 0.f), fmaxf(b23.x, 0.f), fmaxf(b23.y, 0.f)};
      unsigned code = 0;
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const float d = ra[e] - rb[e];
        acc += fabsf(d);
        code |= (fa[e] > 0.f ? (d > 0.f ? 3u : (d < 0.f ? 1u : 2u)) : 0u) << (2 * e);
      }
      if (V16a) {
        const __nv_bfloat162 h0 = __floats2bfloat162_rn(ra[0], ra[1]), h1 = __floats2bfloat162_rn(ra[2], ra[3]);
        *reinterpret_cast<uint2*>(V16a + (p + i) * g.C + c) = make_uint2(*reinterpret_cast<const uint32_t*>(&h0), *reinterpret_cast<const uint32_t*>(&h1));
      }
      if (V16b) {
        const __nv_bfloat162 h0 = __floats2bfloat162_rn(rb[0], rb[1]), h1 = __floats2bfloat162_rn(rb[2], rb[3]);
        *reinterpret_cast<uint2*>(V16b + (p + i) * g.C + c) = make_uint2(*reinterpret_cast<const uint32_t*>(&h0), *reinterpret_cast<const uint32_t*>(&h1));
      }
      if (mask) mask[(p + i) * lanes + lc] = (uint8_t)code;
    }
    if (++cnt == 4) { dacc += (double)acc; acc = 0.f; cnt = 0; }
  }
  dacc += (double)acc;
  __shared__ double sh[32];
  const double tot = block_sum<double>(dacc, sh);
  if (threadIdx.x == 0) partial[blockIdx.x] = tot;
}
// loss[0] (+)= coef * sum_b partial[b]   (one block, fixed order: deterministic)
__global__ void __launch_bounds__(256) l1_fields_finish_kernel(const double* __restrict__ partial, int nblocks, double coef, float* loss, int accumulate) {
  __shared__ double sh[32];
  double a = 0.0;
  for (int b = threadIdx.x; b < nblocks; b += blockDim.x) a += partial[b];
  a = block_sum<double>(a, sh);
  if (threadIdx.x == 0) { const float v = (float)(coef * a); loss[0] = accumulate ? loss[0] + v : v; }
}

// ---------------------------------------------------------------------------------------------------------------- reduce
// One block = one 16 x 32 tile of S.  Phase 1: every V pixel that feeds the tile is read ONCE (8 lanes per pixel, C/8 channels per
// lane as 128-byte coalesced float4 groups); its nine per-tap channel dot products  t[k] = sum_c V[p][c] w[c][k]  are reduced over
// the 8 lanes with a 10-shuffle reduce-scatter and parked in shared memory (zero for pixels outside the image).  Phase 2: each
// output pixel gathers its nine taps.  transposed == 0: C -> 1 forward conv, stride 1: v = q + k - pad.  transposed == 1: data
// gradient of the 1 -> C conv: v = (q + pad - k) / stride when divisible.  Fixed summation order: deterministic.
constexpr int RT_H = 16, RT_W = 32, R_MAXPIX = (RT_H + 2) * (RT_W + 2);

// MASKED (data gradient of a tapped 1 -> C conv + ReLU, stride 1, pad 1: VGG19 conv1_1 with the L1 term at relu1_1): neither the activation output
// (the gate) nor the L1 term's gradient exists as a tensor; tap_pair_kernel left 2 bits per element (one byte per pixel and 4 channels):
//   code 0: ReLU closed, dz = 0;   else dz[p][c] = V[p][c] + (code - 2) * gcoef      (code - 2 = sign(fa' - fb'))
template <int CPL, bool MASKED = false>   // channels per lane: C = 8 * CPL
__global__ void __launch_bounds__(256, CPL <= 8 ? 2 : 1) reduce_kernel(const float* __restrict__ V, int v_pitch, const float* __restrict__ w, const float* __restrict__ bias,
                                                                       float* S, const float* res /* may alias S */, Geo g, int transposed, int tiles_x, int tiles_y,
                                                                       const float* __restrict__ gate, int gate_pitch, float gate_slope,
                                                                       const uint8_t* __restrict__ mask = nullptr, float gcoef = 0.f) {
  __shared__ float Ts[9][R_MAXPIX + 4];
  // MASKED: one code byte covers 4 channels; dz = V * mult + add with (mult, add) = (0, 0) for a closed ReLU, else (1, sign * gcoef): two float4 table
  // entries per byte (the arithmetic decode -- shift, mask, compare, convert per element -- made the kernel instruction bound: 0.71 ms for 1.1 GB)
  __shared__ float4 Lm[MASKED ? 256 : 1], La[MASKED ? 256 : 1];
  if (MASKED) {
    const int cb = threadIdx.x;                // blockDim.x == 256
    float m[4], a[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const int code = (cb >> (2 * e)) & 3;
      m[e] = code ? 1.f : 0.f;
      a[e] = code == 1 ? -gcoef : (code == 3 ? gcoef : 0.f);
    }
    Lm[cb] = make_float4(m[0], m[1], m[2], m[3]);
    La[cb] = make_float4(a[0], a[1], a[2], a[3]);
    __syncthreads();
  }
  constexpr int NJ = CPL / 4;
  const int lane8 = threadIdx.x & 7, slot = threadIdx.x >> 3;   // 32 pixel slots per pass
  float wr[NJ][4][9];
#pragma unroll
  for (int j = 0; j < NJ; ++j)
#pragma unroll
    for (int e = 0; e < 4; ++e)
#pragma unroll
      for (int k = 0; k < 9; ++k) wr[j][e][k] = __ldg(w + (j * 32 + lane8 * 4 + e) * 9 + k);
  int t = blockIdx.x;
  const int tx = t % tiles_x; t /= tiles_x;
  const int ty = t % tiles_y, b = t / tiles_y;
  const int qy0 = ty * RT_H, qx0 = tx * RT_W;
  int vy_lo, vx_lo, vy_n, vx_n;
  if (!transposed) { vy_lo = qy0 - g.pad; vx_lo = qx0 - g.pad; vy_n = RT_H + 2; vx_n = RT_W + 2; }
  else {
    const int ey = qy0 + g.pad - 2, ex = qx0 + g.pad - 2;
    vy_lo = ((ey > 0 ? ey : 0) + g.stride - 1) / g.stride; vx_lo = ((ex > 0 ? ex : 0) + g.stride - 1) / g.stride;
    vy_n = (qy0 + RT_H - 1 + g.pad) / g.stride - vy_lo + 1; vx_n = (qx0 + RT_W - 1 + g.pad) / g.stride - vx_lo + 1;
  }
  const int n = vy_n * vx_n;
  const int passes = (n + 31) >> 5;
  const float* vb = V + (size_t)b * g.Hv * g.Wv * v_pitch + lane8 * 4;
  const float* gb = gate ? gate + (size_t)b * g.Hv * g.Wv * gate_pitch + lane8 * 4 : nullptr;
  // gate: V is the gradient of an activation's OUTPUT and gate that output (ReLU / LeakyReLU): the wide tensor is multiplied by act'(gate) while it is
  // loaded, so that the activation backward is not a pass of its own (1 GB read + 1 GB written per 64-channel 256x512 batch of 32)
  const uint8_t* mb = MASKED ? mask + (size_t)b * g.Hv * g.Wv * (g.C >> 2) + lane8 : nullptr;
  auto load = [&](int i, float4* dst, unsigned& mk) {
    const int ry = i / vx_n, rx = i - ry * vx_n;
    const int vy = vy_lo + ry, vx = vx_lo + rx;
    const bool ok = i < n && vy >= 0 && vy < g.Hv && vx >= 0 && vx < g.Wv;
    const size_t pix = (size_t)(ok ? vy * g.Wv + vx : 0);
    const float* src = vb + pix * v_pitch;
    if (MASKED) {
      mk = 0u;
      if (ok) {
#pragma unroll
        for (int j = 0; j < NJ; ++j) mk |= (unsigned)__ldg(mb + pix * (g.C >> 2) + j * 8) << (8 * j);
      }
    }
#pragma unroll
    for (int j = 0; j < NJ; ++j) {
      dst[j] = ok ? __ldg(reinterpret_cast<const float4*>(src + j * 32)) : make_float4(0.f, 0.f, 0.f, 0.f);
      if (gb && ok) {
        const float4 q = __ldg(reinterpret_cast<const float4*>(gb + pix * gate_pitch + j * 32));
        dst[j].x = q.x > 0.f ? dst[j].x : gate_slope * dst[j].x; dst[j].y = q.y > 0.f ? dst[j].y : gate_slope * dst[j].y;
        dst[j].z = q.z > 0.f ? dst[j].z : gate_slope * dst[j].z; dst[j].w = q.w > 0.f ? dst[j].w : gate_slope * dst[j].w;
      }
    }
  };
  const bool b2 = lane8 & 4, b1 = lane8 & 2, b0 = lane8 & 1;
  float4 cur[NJ], nxt[NJ];
  unsigned mcur = 0u, mnxt = 0u;
  load(slot, cur, mcur);
  for (int ps = 0; ps < passes; ++ps) {
    const int i = ps * 32 + slot;
    load(i + 32, nxt, mnxt);                // next pass in flight while this one is reduced (i + 32 >= n loads nothing)
    if (MASKED) {
#pragma unroll
      for (int j = 0; j < NJ; ++j) {
        const unsigned cb = (mcur >> (8 * j)) & 255u;
        const float4 m = Lm[cb], a = La[cb];
        // V * 1 + (+-gcoef | 0) is the same single fp32 addition as before; a closed gate gives exactly 0 (V is finite)
        cur[j] = make_float4(fmaf(cur[j].x, m.x, a.x), fmaf(cur[j].y, m.y, a.y), fmaf(cur[j].z, m.z, a.z), fmaf(cur[j].w, m.w, a.w));
      }
    }
    // nine per-tap dot products over this lane's channels: taps in pairs on the packed fp32x2 FMA (5 instead of 9 instructions per channel)
    float2 t2[4] = {make_float2(0.f, 0.f), make_float2(0.f, 0.f), make_float2(0.f, 0.f), make_float2(0.f, 0.f)};
    float t8s = 0.f;
#pragma unroll
    for (int j = 0; j < NJ; ++j) {
      const float cv[4] = {cur[j].x, cur[j].y, cur[j].z, cur[j].w};
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const float2 v2 = make_float2(cv[e], cv[e]);
#pragma unroll
        for (int q = 0; q < 4; ++q) t2[q] = __ffma2_rn(v2, make_float2(wr[j][e][2 * q], wr[j][e][2 * q + 1]), t2[q]);
        t8s = fmaf(cv[e], wr[j][e][8], t8s);
      }
    }
    const float tk[9] = {t2[0].x, t2[0].y, t2[1].x, t2[1].y, t2[2].x, t2[2].y, t2[3].x, t2[3].y, t8s};
    // reduce-scatter over the 8 lanes of the pixel: lane l ends up with the full sum of tap l; tap 8 is reduced everywhere
    float u[4], v2[2];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const float send = b2 ? tk[q] : tk[q + 4], keep = b2 ? tk[q + 4] : tk[q];
      u[q] = keep + __shfl_xor_sync(0xffffffffu, send, 4);
    }
    float t8 = tk[8] + __shfl_xor_sync(0xffffffffu, tk[8], 4);
#pragma unroll
    for (int q = 0; q < 2; ++q) {
      const float send = b1 ? u[q] : u[q + 2], keep = b1 ? u[q + 2] : u[q];
      v2[q] = keep + __shfl_xor_sync(0xffffffffu, send, 2);
    }
    t8 += __shfl_xor_sync(0xffffffffu, t8, 2);
    const float send = b0 ? v2[0] : v2[1], keep = b0 ? v2[1] : v2[0];
    const float own = keep + __shfl_xor_sync(0xffffffffu, send, 1);
    t8 += __shfl_xor_sync(0xffffffffu, t8, 1);
    if (i < n) {
      Ts[lane8][i] = own;
      if (lane8 == 0) Ts[8][i] = t8;
    }
#pragma unroll
    for (int j = 0; j < NJ; ++j) cur[j] = nxt[j];
    mcur = mnxt;
  }
  __syncthreads();
  const float bs = bias ? __ldg(bias) : 0.f;
  for (int o = threadIdx.x; o < RT_H * RT_W; o += 256) {
    const int oy = o / RT_W, ox = o % RT_W;
    const int qy = qy0 + oy, qx = qx0 + ox;
    if (qy >= g.Hs || qx >= g.Ws) continue;
    float acc = 0.f;
#pragma unroll
    for (int kh = 0; kh < 3; ++kh) {
      int ry;
      if (!transposed) ry = oy + kh;
      else { const int e = qy + g.pad - kh; if (e < 0 || e % g.stride) continue; ry = e / g.stride - vy_lo; }
#pragma unroll
      for (int kw = 0; kw < 3; ++kw) {
        int rx;
        if (!transposed) rx = ox + kw;
        else { const int e = qx + g.pad - kw; if (e < 0 || e % g.stride) continue; rx = e / g.stride - vx_lo; }
        acc += Ts[kh * 3 + kw][ry * vx_n + rx];
      }
    }
    const size_t q = ((size_t)b * g.Hs + qy) * g.Ws + qx;
    S[q] = acc + bs + (res ? res[q] : 0.f);
  }
}

// ---------------------------------------------------------------------------------------------------------------- wgrad
// partial[block][c][k] = sum over the block's pixel groups of V[p][c] * S[p*stride + k - pad]   (flip: S at p + pad - k)
template <int STRIDE, int PX>
__global__ void __launch_bounds__(256, 2) wgrad_kernel(const float* __restrict__ V, int v_pitch, const float* __restrict__ S, float* __restrict__ partial, Geo g,
                                                       int groups_per_block, int flip) {
  extern __shared__ float sh[];                  // [256 threads][37]
  constexpr int NC = STRIDE ? (PX - 1) * STRIDE + 3 : 3;
  const int lanes = g.C >> 2;
  const int gpp = 256 / lanes;
  const int lane_c = threadIdx.x % lanes, prow = threadIdx.x / lanes;
  const int c = lane_c << 2;
  const int stride = STRIDE ? STRIDE : g.stride;
  const int pad = flip ? 2 - g.pad : g.pad;
  const int gx = g.Wv / PX;
  const int total = g.B * g.Hv * gx;
  const int g0 = blockIdx.x * groups_per_block, g1 = (g0 + groups_per_block < total) ? g0 + groups_per_block : total;
  float acc[36];
#pragma unroll
  for (int i = 0; i < 36; ++i) acc[i] = 0.f;
  for (int gi = g0 + prow; gi < g1; gi += gpp) {
    const int xg = gi % gx, r = gi / gx, y = r % g.Hv, b = r / g.Hv;
    const int x0 = xg * PX;
    const size_t p = ((size_t)b * g.Hv + y) * g.Wv + x0;
    float4 v[PX];
#pragma unroll
    for (int i = 0; i < PX; ++i) v[i] = __ldg(reinterpret_cast<const float4*>(V + (p + i) * v_pitch + c));
    const float* sb = S + (size_t)b * g.Hs * g.Ws;
    const int sx0 = x0 * stride - pad;
    float sw[3][NC];
#pragma unroll
    for (int kh = 0; kh < 3; ++kh) {
      const int sy = y * stride + kh - pad;
      const bool rowok = sy >= 0 && sy < g.Hs;
      const float* sr = sb + (size_t)(rowok ? sy : 0) * g.Ws;
#pragma unroll
      for (int j = 0; j < NC; ++j) {
        const int sx = sx0 + j;
        sw[kh][j] = (rowok && sx >= 0 && sx < g.Ws) ? __ldg(sr + sx) : 0.f;
      }
    }
#pragma unroll
    for (int i = 0; i < PX; ++i)
#pragma unroll
      for (int kh = 0; kh < 3; ++kh)
#pragma unroll
        for (int kw = 0; kw < 3; ++kw) {
          const float s = sw[kh][(STRIDE ? i * STRIDE : 0) + kw];
          const int k = kh * 3 + kw;
          acc[k] = fmaf(v[i].x, s, acc[k]); acc[9 + k] = fmaf(v[i].y, s, acc[9 + k]); acc[18 + k] = fmaf(v[i].z, s, acc[18 + k]); acc[27 + k] = fmaf(v[i].w, s, acc[27 + k]);
        }
  }
#pragma unroll
  for (int i = 0; i < 36; ++i) sh[threadIdx.x * 37 + i] = acc[i];
  __syncthreads();
  // thread t < lanes*36 sums the gpp pixel rows of (lane_c = t / 36, i = t % 36)
  for (int t = threadIdx.x; t < lanes * 36; t += 256) {
    const int lc = t / 36, i = t % 36;
    float s = 0.f;
    for (int rws = 0; rws < gpp; ++rws) s += sh[(rws * lanes + lc) * 37 + i];
    const int k = i % 9;
    partial[(size_t)blockIdx.x * g.C * 9 + ((lc << 2) + i / 9) * 9 + (flip ? 8 - k : k)] = s;
  }
}
// out[i] (+)= sum_b partial[b][i]: one block per output element, double accumulation in a fixed order
__global__ void __launch_bounds__(128) wgrad_final_kernel(const float* __restrict__ partial, int blocks, int n, float* __restrict__ out, int accumulate) {
  __shared__ double red[128];
  const int i = blockIdx.x;
  double s = 0.0;
  for (int b = threadIdx.x; b < blocks; b += 128) s += (double)partial[(size_t)b * n + i];
  red[threadIdx.x] = s;
  __syncthreads();
  for (int o = 64; o > 0; o >>= 1) {
    if (threadIdx.x < o) red[threadIdx.x] += red[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) out[i] = accumulate ? out[i] + (float)red[0] : (float)red[0];
}
}  // namespace thin
}  // namespace gdn


// ------------------------------------------------------------------------------------------------ narrow 1x1 convolutions (C <-> T <= 16 channels)
// The generator head's tap planes (engine.op_upsample_skip_final): z[px][t] = sum_c u[px][c] wp[t][c] and its two gradients.  HBM-bound single passes over
// the C-channel tensor (the fp32 implicit-GEMM engine ran these at 0.8 TB/s: N = 12 is far below its tile).  T <= 16, C % 4 == 0, C <= 256.
namespace gdn {
namespace narrow {
constexpr int NT = 16;
// thread = pixel; weights [T][C] in shared memory (broadcast reads)
__global__ void __launch_bounds__(256) fwd_kernel(const float* __restrict__ u, int C, const float* __restrict__ wp, int T, float* __restrict__ z, long long M) {
  extern __shared__ float ws[];
  for (int i = threadIdx.x; i < T * C; i += blockDim.x) ws[i] = wp[i];
  __syncthreads();
  for (long long px = blockIdx.x * (long long)blockDim.x + threadIdx.x; px < M; px += (long long)gridDim.x * blockDim.x) {
    float acc[NT];
#pragma unroll
    for (int t = 0; t < NT; ++t) acc[t] = 0.f;
    const float4* row = reinterpret_cast<const float4*>(u + (size_t)px * C);
    for (int q = 0; q < (C >> 2); ++q) {
      const float4 v = __ldg(row + q);
#pragma unroll
      for (int t = 0; t < NT; ++t)
        if (t < T) {
          const float4 w4 = *reinterpret_cast<const float4*>(ws + t * C + 4 * q);
          acc[t] = fmaf(v.x, w4.x, fmaf(v.y, w4.y, fmaf(v.z, w4.z, fmaf(v.w, w4.w, acc[t]))));
        }
    }
    float* o = z + (size_t)px * T;
#pragma unroll
    for (int t = 0; t < NT; ++t)
      if (t < T) o[t] = acc[t];
  }
}
// du[px][c] (+)= sum_t dz[px][t] wp[t][c]; thread = (pixel, 4-channel chunk): the C/4 threads of a pixel write one contiguous row
__global__ void __launch_bounds__(256) dgrad_kernel(const float* __restrict__ dz, int T, const float* __restrict__ wp, int C, float* __restrict__ du, long long M, int accumulate) {
  extern __shared__ float ws[];
  for (int i = threadIdx.x; i < T * C; i += blockDim.x) ws[i] = wp[i];
  __syncthreads();
  const int Cv = C >> 2;
  const long long total = M * Cv;
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
    const long long px = idx / Cv; const int q = (int)(idx - px * Cv);
    float4 acc = accumulate ? *reinterpret_cast<const float4*>(du + (size_t)px * C + 4 * q) : make_float4(0.f, 0.f, 0.f, 0.f);
    const float* g = dz + (size_t)px * T;
#pragma unroll
    for (int t = 0; t < NT; ++t)
      if (t < T) {
        const float gv = __ldg(g + t);
        const float4 w4 = *reinterpret_cast<const float4*>(ws + t * C + 4 * q);
        acc.x = fmaf(gv, w4.x, acc.x); acc.y = fmaf(gv, w4.y, acc.y); acc.z = fmaf(gv, w4.z, acc.z); acc.w = fmaf(gv, w4.w, acc.w);
      }
    *reinterpret_cast<float4*>(du + (size_t)px * C + 4 * q) = acc;
  }
}
// partial[block][t][c] = sum over the block's pixels of dz[px][t] u[px][c]; thread = (4-channel chunk, pixel slot); fixed order => deterministic
__global__ void __launch_bounds__(256) wgrad_kernel(const float* __restrict__ dz, int T, const float* __restrict__ u, int C, long long M, long long px_per_block,
                                                    float* __restrict__ partial) {
  extern __shared__ float red[];                 // [slots][T][C]
  const int Cv = C >> 2, slots = blockDim.x / Cv;
  const int q = threadIdx.x % Cv, slot = threadIdx.x / Cv;
  float acc[NT][4];
#pragma unroll
  for (int t = 0; t < NT; ++t) acc[t][0] = acc[t][1] = acc[t][2] = acc[t][3] = 0.f;
  const long long p0 = blockIdx.x * px_per_block, p1 = p0 + px_per_block < M ? p0 + px_per_block : M;
  if (slot < slots)
    for (long long px = p0 + slot; px < p1; px += slots) {
      const float4 v = __ldg(reinterpret_cast<const float4*>(u + (size_t)px * C) + q);
      const float* g = dz + (size_t)px * T;
#pragma unroll
      for (int t = 0; t < NT; ++t)
        if (t < T) {
          const float gv = __ldg(g + t);
          acc[t][0] = fmaf(gv, v.x, acc[t][0]); acc[t][1] = fmaf(gv, v.y, acc[t][1]); acc[t][2] = fmaf(gv, v.z, acc[t][2]); acc[t][3] = fmaf(gv, v.w, acc[t][3]);
        }
    }
  if (slot < slots) {
#pragma unroll
    for (int t = 0; t < NT; ++t)
      if (t < T) *reinterpret_cast<float4*>(red + ((size_t)slot * T + t) * C + 4 * q) = make_float4(acc[t][0], acc[t][1], acc[t][2], acc[t][3]);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < T * C; i += blockDim.x) {
    float sum = 0.f;
    for (int sl = 0; sl < slots; ++sl) sum += red[(size_t)sl * T * C + i];
    partial[(size_t)blockIdx.x * T * C + i] = sum;
  }
}
__global__ void wgrad_final_kernel(const float* __restrict__ partial, int nblocks, int n, float* __restrict__ out, int accumulate) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float s0 = 0.f, s1 = 0.f;
  int b = 0;
  for (; b + 1 < nblocks; b += 2) { s0 += partial[(size_t)b * n + i]; s1 += partial[(size_t)(b + 1) * n + i]; }
  if (b < nblocks) s0 += partial[(size_t)b * n + i];
  out[i] = accumulate ? out[i] + (s0 + s1) : s0 + s1;
}
}  // namespace narrow
}  // namespace gdn

using namespace gdn;
using namespace gdn::thin;

static bool thin_ok(int C) { return C == 32 || C == 64 || C == 128; }

extern "C" int gdn_thin_conv_supported(int C, int kh, int kw) { return thin_ok(C) && kh == 3 && kw == 3; }

extern "C" int gdn_thin_conv_expand_p(const float* s_in, const float* w, const float* bias, float* v_out, int v_pitch, const float* res, int res_pitch,
                                      int B, int Hv, int Wv, int C, int Hs, int Ws, int stride, int pad, int flip, int act, float slope, uint16_t* v16, gdn_stream_t st);
extern "C" int gdn_thin_conv_expand(const float* s_in, const float* w, const float* bias, float* v_out, int v_pitch, const float* res, int res_pitch,
                                    int B, int Hv, int Wv, int C, int Hs, int Ws, int stride, int pad, int flip, int act, float slope, gdn_stream_t st) {
  return gdn_thin_conv_expand_p(s_in, w, bias, v_out, v_pitch, res, res_pitch, B, Hv, Wv, C, Hs, Ws, stride, pad, flip, act, slope, nullptr, st);
}
extern "C" int gdn_thin_conv_expand_p(const float* s_in, const float* w, const float* bias, float* v_out, int v_pitch, const float* res, int res_pitch,
                                      int B, int Hv, int Wv, int C, int Hs, int Ws, int stride, int pad, int flip, int act, float slope, uint16_t* v16, gdn_stream_t st) {
  GDN_CHECK_ARG(((uintptr_t)v16 & 7) == 0);
  __nv_bfloat16* V16 = reinterpret_cast<__nv_bfloat16*>(v16);
  GDN_CHECK_ARG(s_in && w && (v_out || (v16 && !res)) && thin_ok(C) && (!v_out || (v_pitch >= C && v_pitch % 4 == 0 && ((uintptr_t)v_out & 15) == 0)) && (!bias || ((uintptr_t)bias & 15) == 0));
  GDN_CHECK_ARG(!res || (res_pitch % 4 == 0 && ((uintptr_t)res & 15) == 0));
  GDN_CHECK_ARG((long long)B * Hv * Wv < (1ll << 31) && stride >= 1);
  Geo g = {B, Hv, Wv, C, Hs, Ws, stride, pad};
  const int gpp = 256 / (C / 4);
  const bool px4 = Wv % 4 == 0 && (stride == 1 || stride == 2);
  const long long groups = (long long)B * Hv * (px4 ? Wv / 4 : Wv);
  long long blocks = cdiv(groups, gpp);
  if (blocks > 8 * kNumSMs) blocks = 8 * kNumSMs;
  if (blocks < 1) blocks = 1;
  cudaStream_t s = as_stream(st);
  if (px4 && stride == 1) expand_kernel<1, 4><<<(unsigned)blocks, 256, 0, s>>>(s_in, w, bias, v_out, v_pitch, res, res_pitch, g, flip, act, slope, V16);
  else if (px4) expand_kernel<2, 4><<<(unsigned)blocks, 256, 0, s>>>(s_in, w, bias, v_out, v_pitch, res, res_pitch, g, flip, act, slope, V16);
  else expand_kernel<0, 1><<<(unsigned)blocks, 256, 0, s>>>(s_in, w, bias, v_out, v_pitch, res, res_pitch, g, flip, act, slope, V16);
  GDN_CHECK_LAUNCH();
  return GDN_OK;
}

extern "C" int gdn_thin_conv_reduce_gated(const float* v_in, int v_pitch, const float* gate, int gate_pitch, float gate_slope, const float* w, const float* bias, float* s_out,
                                          const float* res, int B, int Hv, int Wv, int C, int Hs, int Ws, int stride, int pad, int transposed, gdn_stream_t st);
extern "C" int gdn_thin_conv_reduce(const float* v_in, int v_pitch, const float* w, const float* bias, float* s_out, const float* res,
                                    int B, int Hv, int Wv, int C, int Hs, int Ws, int stride, int pad, int transposed, gdn_stream_t st) {
  return gdn_thin_conv_reduce_gated(v_in, v_pitch, nullptr, 0, 0.f, w, bias, s_out, res, B, Hv, Wv, C, Hs, Ws, stride, pad, transposed, st);
}
extern "C" int gdn_thin_conv_reduce_gated(const float* v_in, int v_pitch, const float* gate, int gate_pitch, float gate_slope, const float* w, const float* bias, float* s_out,
                                          const float* res, int B, int Hv, int Wv, int C, int Hs, int Ws, int stride, int pad, int transposed, gdn_stream_t st) {
  GDN_CHECK_ARG(!gate || (gate_pitch >= C && gate_pitch % 4 == 0 && ((uintptr_t)gate & 15) == 0));
  GDN_CHECK_ARG(v_in && w && s_out && thin_ok(C) && v_pitch >= C && v_pitch % 4 == 0 && ((uintptr_t)v_in & 15) == 0 && (transposed || stride == 1) && stride >= 1);
  GDN_CHECK_ARG((long long)B * Hv * Wv < (1ll << 31) && pad >= 0 && pad <= 2);
  Geo g = {B, Hv, Wv, C, Hs, Ws, stride, pad};
  const int tiles_x = (int)cdiv(Ws, RT_W), tiles_y = (int)cdiv(Hs, RT_H);
  const long long blocks = (long long)B * tiles_x * tiles_y;
  GDN_CHECK_ARG(blocks < (1ll << 31));
  cudaStream_t s = as_stream(st);
  if (C == 32) reduce_kernel<4><<<(unsigned)blocks, 256, 0, s>>>(v_in, v_pitch, w, bias, s_out, res, g, transposed, tiles_x, tiles_y, gate, gate_pitch, gate_slope);
  else if (C == 64) reduce_kernel<8><<<(unsigned)blocks, 256, 0, s>>>(v_in, v_pitch, w, bias, s_out, res, g, transposed, tiles_x, tiles_y, gate, gate_pitch, gate_slope);
  else reduce_kernel<16><<<(unsigned)blocks, 256, 0, s>>>(v_in, v_pitch, w, bias, s_out, res, g, transposed, tiles_x, tiles_y, gate, gate_pitch, gate_slope);
  GDN_CHECK_LAUNCH();
  return GDN_OK;
}

extern "C" int gdn_thin_conv_tap_supported(int C, int H, int W) { return thin_ok(C) && C <= 64 && W % 4 == 0 && H > 0; }
extern "C" size_t gdn_thin_conv_tap_mask_bytes(int B, int H, int W, int C) { return (size_t)B * H * W * (C / 4); }
extern "C" int gdn_thin_conv_tap_pair(const float* fa, const float* fb, const float* w, const float* cbias, int B, int H, int W, int C, float* loss, int loss_accumulate,
                                      float scale, uint16_t* v16a, uint16_t* v16b, uint8_t* mask, void* ws, size_t ws_bytes, gdn_stream_t st) {
  GDN_CHECK_ARG(fa && fb && w && loss && ws && gdn_thin_conv_tap_supported(C, H, W) && (!cbias || ((uintptr_t)cbias & 15) == 0));
  GDN_CHECK_ARG((long long)B * H * W < (1ll << 31) && ((uintptr_t)v16a & 7) == 0 && ((uintptr_t)v16b & 7) == 0);
  Geo g = {B, H, W, C, H, W, 1, 1};
  const int gpp = 256 / (C / 4);
  long long blocks = cdiv((long long)B * H * (W / 4), gpp);
  if (blocks > 8 * kNumSMs) blocks = 8 * kNumSMs;
  if (ws_bytes < (size_t)blocks * sizeof(double)) { set_error("gdn_thin_conv_tap_pair: workspace too small"); return GDN_EWORKSPACE; }
  cudaStream_t s = as_stream(st);
  tap_pair_kernel<4><<<(unsigned)blocks, 256, 0, s>>>(fa, fb, w, cbias, g, reinterpret_cast<double*>(ws), reinterpret_cast<__nv_bfloat16*>(v16a),
                                                       reinterpret_cast<__nv_bfloat16*>(v16b), mask);
  GDN_CHECK_LAUNCH();
  l1_fields_finish_kernel<<<1, 256, 0, s>>>(reinterpret_cast<const double*>(ws), (int)blocks, (double)scale / ((double)B * H * W * C), loss, loss_accumulate);
  GDN_CHECK_LAUNCH();
  return GDN_OK;
}
extern "C" int gdn_thin_conv_tap_dgrad(const float* dy, int dy_pitch, const uint8_t* mask, const float* w, float gcoef, float* s_out, const float* res,
                                       int B, int H, int W, int C, gdn_stream_t st) {
  GDN_CHECK_ARG(dy && mask && w && s_out && gdn_thin_conv_tap_supported(C, H, W) && dy_pitch >= C && dy_pitch % 4 == 0 && ((uintptr_t)dy & 15) == 0);
  GDN_CHECK_ARG((long long)B * H * W < (1ll << 31));
  Geo g = {B, H, W, C, H, W, 1, 1};
  const int tiles_x = (int)cdiv(W, RT_W), tiles_y = (int)cdiv(H, RT_H);
  const long long blocks = (long long)B * tiles_x * tiles_y;
  GDN_CHECK_ARG(blocks < (1ll << 31));
  cudaStream_t s = as_stream(st);
  if (C == 32) reduce_kernel<4, true><<<(unsigned)blocks, 256, 0, s>>>(dy, dy_pitch, w, nullptr, s_out, res, g, 1, tiles_x, tiles_y, nullptr, 0, 0.f, mask, gcoef);
  else reduce_kernel<8, true><<<(unsigned)blocks, 256, 0, s>>>(dy, dy_pitch, w, nullptr, s_out, res, g, 1, tiles_x, tiles_y, nullptr, 0, 0.f, mask, gcoef);
  GDN_CHECK_LAUNCH();
  return GDN_OK;
}

static int thin_wgrad_blocks(int B, int Hv, int Wv, int C) {
  const long long groups = (long long)B * Hv * (Wv % 4 == 0 ? Wv / 4 : Wv);
  long long b = cdiv(groups, (256 / (C / 4)) * 8);
  return (int)(b < 4 * kNumSMs ? (b > 0 ? b : 1) : 4 * kNumSMs);
}
extern "C" size_t gdn_thin_conv_wgrad_ws_bytes(int B, int Hv, int Wv, int C) {
  return thin_ok(C) ? (size_t)thin_wgrad_blocks(B, Hv, Wv, C) * C * 9 * sizeof(float) : 0;
}

extern "C" int gdn_thin_conv_wgrad(const float* v, int v_pitch, const float* s_in, float* dw, int accumulate, int B, int Hv, int Wv, int C, int Hs, int Ws, int stride, int pad,
                                   int flip, float* ws, size_t ws_bytes, gdn_stream_t st) {
  GDN_CHECK_ARG(v && s_in && dw && ws && thin_ok(C) && v_pitch >= C && v_pitch % 4 == 0 && ((uintptr_t)v & 15) == 0 && (!flip || stride == 1) && stride >= 1);
  GDN_CHECK_ARG((long long)B * Hv * Wv < (1ll << 31));
  if (ws_bytes < gdn_thin_conv_wgrad_ws_bytes(B, Hv, Wv, C)) { set_error("gdn_thin_conv_wgrad: workspace too small"); return GDN_EWORKSPACE; }
  const bool px4 = Wv % 4 == 0 && (stride == 1 || stride == 2);
  const long long groups_run = (long long)B * Hv * (px4 ? Wv / 4 : Wv);
  const int blocks = thin_wgrad_blocks(B, Hv, Wv, C);
  const int gpb = (int)cdiv(groups_run, blocks);
  Geo g = {B, Hv, Wv, C, Hs, Ws, stride, pad};
  cudaStream_t s = as_stream(st);
  const size_t smem = 256 * 37 * sizeof(float);
  if (px4 && stride == 1) wgrad_kernel<1, 4><<<blocks, 256, smem, s>>>(v, v_pitch, s_in, ws, g, gpb, flip);
  else if (px4) wgrad_kernel<2, 4><<<blocks, 256, smem, s>>>(v, v_pitch, s_in, ws, g, gpb, flip);
  else wgrad_kernel<0, 1><<<blocks, 256, smem, s>>>(v, v_pitch, s_in, ws, g, gpb, flip);
  GDN_CHECK_LAUNCH();
  wgrad_final_kernel<<<C * 9, 128, 0, s>>>(ws, blocks, C * 9, dw, accumulate);
  GDN_CHECK_LAUNCH();
  return GDN_OK;
}

// ---- narrow 1x1 convolutions
static inline bool narrow_ok(int C, int T) { return C > 0 && C % 4 == 0 && C <= 256 && T > 0 && T <= gdn::narrow::NT; }
static inline int narrow_wgrad_blocks(long long M) { long long b = cdiv(M, 512); return (int)(b < 4 * kNumSMs ? (b > 0 ? b : 1) : 4 * kNumSMs); }
extern "C" int gdn_narrow_conv1x1_fwd(const float* u, int C, const float* wp, int T, float* z, long long M, gdn_stream_t st) {
  GDN_CHECK_ARG(u && wp && z && M > 0 && narrow_ok(C, T) && ((uintptr_t)u & 15) == 0);
  const long long blocks = cdiv(M, 256);
  gdn::narrow::fwd_kernel<<<(unsigned)(blocks < 16 * kNumSMs ? blocks : 16 * kNumSMs), 256, (size_t)T * C * sizeof(float), as_stream(st)>>>(u, C, wp, T, z, M);
  GDN_CHECK_LAUNCH();
  return GDN_OK;
}
extern "C" int gdn_narrow_conv1x1_dgrad(const float* dz, int T, const float* wp, int C, float* du, long long M, int accumulate, gdn_stream_t st) {
  GDN_CHECK_ARG(dz && wp && du && M > 0 && narrow_ok(C, T) && ((uintptr_t)du & 15) == 0);
  const long long blocks = cdiv(M * (C / 4), 256);
  gdn::narrow::dgrad_kernel<<<(unsigned)(blocks < 16 * kNumSMs ? blocks : 16 * kNumSMs), 256, (size_t)T * C * sizeof(float), as_stream(st)>>>(dz, T, wp, C, du, M, accumulate);
  GDN_CHECK_LAUNCH();
  return GDN_OK;
}
extern "C" size_t gdn_narrow_conv1x1_wgrad_ws_bytes(long long M, int C, int T) { return (size_t)narrow_wgrad_blocks(M) * T * C * sizeof(float); }
extern "C" int gdn_narrow_conv1x1_wgrad(const float* dz, int T, const float* u, int C, long long M, float* gwp, int accumulate, void* ws, size_t ws_bytes, gdn_stream_t st) {
  GDN_CHECK_ARG(dz && u && gwp && ws && M > 0 && narrow_ok(C, T) && ((uintptr_t)u & 15) == 0 && 256 % (C / 4) == 0);
  if (ws_bytes < gdn_narrow_conv1x1_wgrad_ws_bytes(M, C, T)) { set_error("gdn_narrow_conv1x1_wgrad: workspace too small"); return GDN_EWORKSPACE; }
  const int blocks = narrow_wgrad_blocks(M);
  const long long ppb = cdiv(M, blocks);
  const int slots = 256 / (C / 4);
  const size_t smem = (size_t)slots * T * C * sizeof(float);
  GDN_CHECK_ARG(smem <= 200 * 1024);
  if (smem > 48 * 1024) GDN_CHECK_CUDA(cudaFuncSetAttribute(gdn::narrow::wgrad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));   // per device, cheap
  cudaStream_t s = as_stream(st);
  gdn::narrow::wgrad_kernel<<<blocks, 256, smem, s>>>(dz, T, u, C, M, ppb, reinterpret_cast<float*>(ws));
  GDN_CHECK_LAUNCH();
  gdn::narrow::wgrad_final_kernel<<<(unsigned)cdiv(T * C, 128), 128, 0, s>>>(reinterpret_cast<const float*>(ws), blocks, T * C, gwp, accumulate);
  GDN_CHECK_LAUNCH();
  return GDN_OK;
}
