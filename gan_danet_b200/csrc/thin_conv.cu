// "Thin" 3x3 convolutions: one side has a single channel, the other C channels (C % 4 == 0, C <= 128).  They are HBM-bound
// (D conv1 1->64, VGG19 conv1_1 on the channel-summed weight, the generator's final 64->1 conv, and their
// gradients), so they run as coalesced CUDA-core kernels that read the wide tensor exactly once instead of as padded tensor-core tiles.
// Call sites in the reference: models/discriminator.py:62 (conv1), models/generator.py:228 (final), models/losses.py:58,64-65 (conv1_1
// on x.repeat(1,3,1,1)).  All three kernels share one geometry: a "wide" NHWC tensor V [B,Hv,Wv,C] and a single-channel field S
// [B,Hs,Ws], related by  s-coordinate = v-coordinate * stride + (k - pad)  for tap k (forward conv with S as input and V as output,
// or its transpose).
//   expand : V[p][c]  = act(sum_k S[p*stride + k - pad] * w[c][k] + bias[c]) (+ res)         1 -> C forward; data gradient of C -> 1
//   reduce : S[q]     = sum_k sum_c V[p(q,k)][c] * w[c][k] (+ bias) (+ res)                  C -> 1 forward; data gradient of 1 -> C
//   wgrad  : dw[c][k] = sum_p V[p][c] * S[p*stride + k - pad]                                weight gradient of both
// Threads: C/4 lanes per pixel, one float4 of channels each (a warp covers 128/C... pixels x 128..512 contiguous bytes).
#include "common.cuh"

namespace gdn {
namespace thin {

struct Geo { int B, Hv, Wv, C, Hs, Ws, stride, pad; };

// V = act(conv(S, w) + bias) + res.  w: [C][9] (row-major taps kh*3+kw).  flip: use tap (2-kh, 2-kw) offsets (data gradient of C -> 1)
__global__ void __launch_bounds__(256) expand_kernel(const float* __restrict__ S, const float* __restrict__ w, const float* __restrict__ bias, float* V, int v_pitch,
                                                     const float* res, int res_pitch,   /* res may alias V */ Geo g, int flip, int act, float slope) {
  const int lanes = g.C >> 2;
  const long long total = (long long)g.B * g.Hv * g.Wv * lanes;
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(idx % lanes) << 2; long long p = idx / lanes;
    const int x = (int)(p % g.Wv); long long r = p / g.Wv; const int y = (int)(r % g.Hv); const int b = (int)(r / g.Hv);
    float4 acc = bias ? *reinterpret_cast<const float4*>(bias + c) : make_float4(0.f, 0.f, 0.f, 0.f);
    const float* sb = S + (size_t)b * g.Hs * g.Ws;
#pragma unroll
    for (int kh = 0; kh < 3; ++kh) {
      const int sy = y * g.stride + (flip ? g.pad - kh : kh - g.pad);
      if (sy < 0 || sy >= g.Hs) continue;
#pragma unroll
      for (int kw = 0; kw < 3; ++kw) {
        const int sx = x * g.stride + (flip ? g.pad - kw : kw - g.pad);
        if (sx < 0 || sx >= g.Ws) continue;
        const float s = __ldg(sb + (size_t)sy * g.Ws + sx);
        const int k = kh * 3 + kw;
        acc.x = fmaf(s, __ldg(w + (c + 0) * 9 + k), acc.x); acc.y = fmaf(s, __ldg(w + (c + 1) * 9 + k), acc.y);
        acc.z = fmaf(s, __ldg(w + (c + 2) * 9 + k), acc.z); acc.w = fmaf(s, __ldg(w + (c + 3) * 9 + k), acc.w);
      }
    }
    acc.x = apply_act(acc.x, act, slope); acc.y = apply_act(acc.y, act, slope); acc.z = apply_act(acc.z, act, slope); acc.w = apply_act(acc.w, act, slope);
    if (res) { const float4 rr = *reinterpret_cast<const float4*>(res + (size_t)p * res_pitch + c); acc.x += rr.x; acc.y += rr.y; acc.z += rr.z; acc.w += rr.w; }
    *reinterpret_cast<float4*>(V + (size_t)p * v_pitch + c) = acc;
  }
}

// S[q] = sum over taps and channels of V * w (+ bias) (+ res).  transposed == 0: C -> 1 forward conv (V is the input, stride 1 only):
// v = q + k - pad.  transposed == 1: data gradient of the 1 -> C conv: v = (q + pad - k) / stride when divisible.
__global__ void __launch_bounds__(256) reduce_kernel(const float* __restrict__ V, int v_pitch, const float* __restrict__ w, const float* __restrict__ bias, float* S,
                                                     const float* res, Geo g, int transposed) {   /* res may alias S */
  const int lanes = g.C >> 2;                    // power of two <= 32 (checked on the host)
  const long long total = (long long)g.B * g.Hs * g.Ws * lanes;
  const long long stride_all = (long long)gridDim.x * blockDim.x;     // a multiple of 32, so lane groups stay intact
  // whole warps iterate together (the lane-group reduction below uses full-warp shuffles)
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx - (threadIdx.x & 31) < total; idx += stride_all) {
    const bool live = idx < total;
    const long long ii = live ? idx : total - 1;
    const int c = (int)(ii % lanes) << 2; long long q = ii / lanes;
    const int x = (int)(q % g.Ws); long long r = q / g.Ws; const int y = (int)(r % g.Hs); const int b = (int)(r / g.Hs);
    float acc = 0.f;
#pragma unroll
    for (int kh = 0; kh < 3; ++kh) {
      int vy;
      if (!transposed) vy = y + kh - g.pad;
      else { const int e = y + g.pad - kh; if (e < 0 || e % g.stride) continue; vy = e / g.stride; }
      if (vy < 0 || vy >= g.Hv) continue;
#pragma unroll
      for (int kw = 0; kw < 3; ++kw) {
        int vx;
        if (!transposed) vx = x + kw - g.pad;
        else { const int e = x + g.pad - kw; if (e < 0 || e % g.stride) continue; vx = e / g.stride; }
        if (vx < 0 || vx >= g.Wv) continue;
        const float4 v = __ldg(reinterpret_cast<const float4*>(V + (((size_t)b * g.Hv + vy) * g.Wv + vx) * v_pitch + c));
        const int k = kh * 3 + kw;
        acc = fmaf(v.x, __ldg(w + (c + 0) * 9 + k), acc); acc = fmaf(v.y, __ldg(w + (c + 1) * 9 + k), acc);
        acc = fmaf(v.z, __ldg(w + (c + 2) * 9 + k), acc); acc = fmaf(v.w, __ldg(w + (c + 3) * 9 + k), acc);
      }
    }
    for (int o = lanes >> 1; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);      // fixed tree: deterministic
    if (live && (idx % lanes) == 0) S[q] = acc + (bias ? __ldg(bias) : 0.f) + (res ? res[q] : 0.f);
  }
}

// partial[block][c][k] = sum over the block's pixels p of V[p][c] * S[p*stride + k - pad]
// flip: S is read at p + pad - k (weight gradient of the C -> 1 convolution, stride 1)
__global__ void __launch_bounds__(256) wgrad_kernel(const float* __restrict__ V, int v_pitch, const float* __restrict__ S, float* __restrict__ partial, Geo g, long long px_per_block, int flip) {
  extern __shared__ float sh[];                  // [256 threads][36] then reduced
  const int lanes = g.C >> 2;
  const int ppb = 256 / lanes;                   // pixels handled per pass by the block
  const int lane_c = threadIdx.x % lanes, prow = threadIdx.x / lanes;
  const int c = lane_c << 2;
  const long long P = (long long)g.B * g.Hv * g.Wv;
  const long long p0 = blockIdx.x * px_per_block, p1 = (p0 + px_per_block < P) ? p0 + px_per_block : P;
  float acc[36];
#pragma unroll
  for (int i = 0; i < 36; ++i) acc[i] = 0.f;
  if (prow < ppb) {
    for (long long p = p0 + prow; p < p1; p += ppb) {
      const int x = (int)(p % g.Wv); long long r = p / g.Wv; const int y = (int)(r % g.Hv); const int b = (int)(r / g.Hv);
      const float4 v = __ldg(reinterpret_cast<const float4*>(V + (size_t)p * v_pitch + c));
      const float* sb = S + (size_t)b * g.Hs * g.Ws;
#pragma unroll
      for (int kh = 0; kh < 3; ++kh) {
        const int sy = y * g.stride + (flip ? g.pad - kh : kh - g.pad);
#pragma unroll
        for (int kw = 0; kw < 3; ++kw) {
          const int sx = x * g.stride + (flip ? g.pad - kw : kw - g.pad);
          const float s = (sy >= 0 && sy < g.Hs && sx >= 0 && sx < g.Ws) ? __ldg(sb + (size_t)sy * g.Ws + sx) : 0.f;
          const int k = kh * 3 + kw;
          acc[k] = fmaf(v.x, s, acc[k]); acc[9 + k] = fmaf(v.y, s, acc[9 + k]); acc[18 + k] = fmaf(v.z, s, acc[18 + k]); acc[27 + k] = fmaf(v.w, s, acc[27 + k]);
        }
      }
    }
  }
#pragma unroll
  for (int i = 0; i < 36; ++i) sh[threadIdx.x * 37 + i] = acc[i];
  __syncthreads();
  // thread t < lanes*36 sums the ppb pixel rows of (lane_c = t / 36, i = t % 36)
  for (int t = threadIdx.x; t < lanes * 36; t += 256) {
    const int lc = t / 36, i = t % 36;
    float s = 0.f;
    for (int rws = 0; rws < ppb; ++rws) s += sh[(rws * lanes + lc) * 37 + i];
    partial[(size_t)blockIdx.x * g.C * 9 + ((lc << 2) + i / 9) * 9 + (i % 9)] = s;
  }
}
// out[c][k] (+)= scale * sum_b partial[b][c][k]  (double accumulation, fixed order)
__global__ void wgrad_final_kernel(const float* __restrict__ partial, int blocks, int n, float* __restrict__ out, int accumulate) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  double s = 0.0;
  for (int b = 0; b < blocks; ++b) s += (double)partial[(size_t)b * n + i];
  out[i] = accumulate ? out[i] + (float)s : (float)s;
}
}  // namespace thin
}  // namespace gdn

using namespace gdn;
using namespace gdn::thin;

static bool thin_ok(int C) { return C >= 4 && C <= 128 && (C & (C - 1)) == 0; }    // C/4 lanes per pixel: a power of two <= 32
static int thin_grid(long long total) { long long b = cdiv(total, 256); return (int)(b < 16 * kNumSMs ? (b > 0 ? b : 1) : 16 * kNumSMs); }

extern "C" int gdn_thin_conv_supported(int C, int kh, int kw) { return thin_ok(C) && kh == 3 && kw == 3; }

extern "C" int gdn_thin_conv_expand(const float* s_in, const float* w, const float* bias, float* v_out, int v_pitch, const float* res, int res_pitch,
                                    int B, int Hv, int Wv, int C, int Hs, int Ws, int stride, int pad, int flip, int act, float slope, gdn_stream_t st) {
  GDN_CHECK_ARG(s_in && w && v_out && thin_ok(C) && v_pitch >= C && v_pitch % 4 == 0 && ((uintptr_t)v_out & 15) == 0 && (!bias || ((uintptr_t)bias & 15) == 0));
  GDN_CHECK_ARG(!res || (res_pitch % 4 == 0 && ((uintptr_t)res & 15) == 0));
  Geo g = {B, Hv, Wv, C, Hs, Ws, stride, pad};
  expand_kernel<<<thin_grid((long long)B * Hv * Wv * (C / 4)), 256, 0, as_stream(st)>>>(s_in, w, bias, v_out, v_pitch, res, res_pitch, g, flip, act, slope);
  GDN_CHECK_LAUNCH();
  return GDN_OK;
}

extern "C" int gdn_thin_conv_reduce(const float* v_in, int v_pitch, const float* w, const float* bias, float* s_out, const float* res,
                                    int B, int Hv, int Wv, int C, int Hs, int Ws, int stride, int pad, int transposed, gdn_stream_t st) {
  GDN_CHECK_ARG(v_in && w && s_out && thin_ok(C) && v_pitch >= C && v_pitch % 4 == 0 && ((uintptr_t)v_in & 15) == 0 && (transposed || stride == 1));
  Geo g = {B, Hv, Wv, C, Hs, Ws, stride, pad};
  reduce_kernel<<<thin_grid((long long)B * Hs * Ws * (C / 4)), 256, 0, as_stream(st)>>>(v_in, v_pitch, w, bias, s_out, res, g, transposed);
  GDN_CHECK_LAUNCH();
  return GDN_OK;
}

static int thin_wgrad_blocks(long long P) { long long b = cdiv(P, 2048); return (int)(b < 8 * kNumSMs ? (b > 0 ? b : 1) : 8 * kNumSMs); }
extern "C" size_t gdn_thin_conv_wgrad_ws_bytes(int B, int Hv, int Wv, int C) { return (size_t)thin_wgrad_blocks((long long)B * Hv * Wv) * C * 9 * sizeof(float); }

extern "C" int gdn_thin_conv_wgrad(const float* v, int v_pitch, const float* s_in, float* dw, int accumulate, int B, int Hv, int Wv, int C, int Hs, int Ws, int stride, int pad,
                                   int flip, float* ws, size_t ws_bytes, gdn_stream_t st) {
  GDN_CHECK_ARG(v && s_in && dw && ws && thin_ok(C) && v_pitch >= C && v_pitch % 4 == 0 && ((uintptr_t)v & 15) == 0 && (!flip || stride == 1));
  if (ws_bytes < gdn_thin_conv_wgrad_ws_bytes(B, Hv, Wv, C)) { set_error("gdn_thin_conv_wgrad: workspace too small"); return GDN_EWORKSPACE; }
  const long long P = (long long)B * Hv * Wv;
  const int blocks = thin_wgrad_blocks(P);
  Geo g = {B, Hv, Wv, C, Hs, Ws, stride, pad};
  wgrad_kernel<<<blocks, 256, 256 * 37 * sizeof(float), as_stream(st)>>>(v, v_pitch, s_in, ws, g, cdiv(P, blocks), flip);
  GDN_CHECK_LAUNCH();
  wgrad_final_kernel<<<(unsigned)cdiv(C * 9, 128), 128, 0, as_stream(st)>>>(ws, blocks, C * 9, dw, accumulate);
  GDN_CHECK_LAUNCH();
  return GDN_OK;
}
