// HBM-bound NHWC kernels: per-channel statistics, train-mode BatchNorm forward/backward, activations, axpy.
// Reference call sites: nn.BatchNorm2d + nn.ReLU at /root/reference/models/generator.py:32-33,61-62,149-150,
// 189-190,219-220,223-224; nn.LeakyReLU(0.2) at models/discriminator.py:68; nn.ReLU of VGG19 (models/losses.py:58).
// All reductions are deterministic two-stage sums accumulated in double.
#include <cuda_bf16.h>
#include "common.cuh"

namespace gdn {

constexpr int kMaxStatBlocks = 4 * kNumSMs;

struct StatPlan { int blocks; long long rows_per_block; };
static StatPlan stat_plan(long long M) {
  StatPlan p;
  long long rpb = cdiv(M, kMaxStatBlocks);
  if (rpb < 64) rpb = 64;
  p.rows_per_block = rpb;
  p.blocks = (int)cdiv(M, rpb);
  return p;
}


// ---- single-launch statistics (round 2): the blocks of a colreduce launch finish the reduction themselves instead of handing their partial
// rows to colreduce_final_kernel + bn_finalize_kernel (+ sums_to_float_kernel): 70 BatchNorm passes per step were three to four launches each,
// the follow-ups latency-bound (14 + 6 + 4 us).  Two-level "last block done" scheme, fixed membership and fixed summation order at both levels
// => still bitwise deterministic: blocks are grouped by kStatGroup; the last block of a group to arrive (ticket counter) adds the group's
// partial rows in block order into a group row; the last group to finish adds the group rows in group order and finalises.  The counters
// are zero on entry and restored to zero by the finishing block (safe under CUDA-graph replay).
constexpr int kStatGroup = 16;
struct StatTail {
  unsigned* counters;        // [1 + groups], zero on entry; nullptr: no tail (two-launch path)
  double* group_partial;     // [groups][2C]
  double* out;               // [2C] final sums (may be nullptr)
  float* fout;               // [2C] the same as float (BatchNorm backward: dbias | dweight), may be nullptr
  int bn;                    // != 0: BatchNorm2d train-mode finalize
  long long M;
  const float* weight; const float* bias; float eps, momentum;
  float* running_mean; float* running_var; long long* num_batches_tracked;
  float* mean; float* invstd; float* scale; float* shift;
};

__device__ __forceinline__ void bn_finalize_channel(const StatTail& t, int c, double s0, double s1) {
  const double m = s0 / (double)t.M;
  double var = s1 / (double)t.M - m * m;
  if (var < 0.0) var = 0.0;
  const float is = (float)(1.0 / sqrt(var + (double)t.eps));
  const float w = t.weight ? t.weight[c] : 1.f, b = t.bias ? t.bias[c] : 0.f;
  const float sc = w * is;
  t.mean[c] = (float)m; t.invstd[c] = is; t.scale[c] = sc; t.shift[c] = b - (float)m * sc;
  if (t.running_mean) t.running_mean[c] = (1.f - t.momentum) * t.running_mean[c] + t.momentum * (float)m;
  if (t.running_var) {
    const double unbiased = t.M > 1 ? var * ((double)t.M / (double)(t.M - 1)) : var;
    t.running_var[c] = (1.f - t.momentum) * t.running_var[c] + t.momentum * (float)unbiased;
  }
}

// Called by every thread of every block after the block's partial row has been written.
__device__ void stat_tail(const StatTail& t, const double* partial, int C) {
  if (t.counters == nullptr) return;
  __shared__ unsigned s_last;
  const int n2 = 2 * C, nblocks = gridDim.x;
  const int groups = (nblocks + kStatGroup - 1) / kStatGroup, g = blockIdx.x / kStatGroup;
  const int b0 = g * kStatGroup, b1 = (b0 + kStatGroup < nblocks) ? b0 + kStatGroup : nblocks;
  __threadfence();                       // this thread's partial sums are visible device-wide before the ticket is taken
  __syncthreads();
  if (threadIdx.x == 0) s_last = (atomicAdd(&t.counters[1 + g], 1u) == (unsigned)(b1 - b0 - 1)) ? 1u : 0u;
  __syncthreads();
  if (!s_last) return;
  __threadfence();
  for (int j = threadIdx.x; j < n2; j += blockDim.x) {
    double a = 0.0;
    for (int b = b0; b < b1; ++b) a += __ldcg(partial + (size_t)b * n2 + j);
    t.group_partial[(size_t)g * n2 + j] = a;
  }
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) s_last = (atomicAdd(&t.counters[0], 1u) == (unsigned)(groups - 1)) ? 1u : 0u;
  __syncthreads();
  if (!s_last) return;
  __threadfence();
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    double s0 = 0.0, s1 = 0.0;
    for (int q = 0; q < groups; ++q) { s0 += __ldcg(t.group_partial + (size_t)q * n2 + c); s1 += __ldcg(t.group_partial + (size_t)q * n2 + C + c); }
    if (t.out) { t.out[c] = s0; t.out[C + c] = s1; }
    if (t.fout) { t.fout[c] = (float)s0; t.fout[C + c] = (float)s1; }
    if (t.bn) bn_finalize_channel(t, c, s0, s1);
  }
  if (threadIdx.x == 0 && t.bn && t.num_batches_tracked) *t.num_batches_tracked += 1;
  for (int i = threadIdx.x; i <= groups; i += blockDim.x) t.counters[i] = 0u;
}

// MODE 0: (sum x, sum x^2).  MODE 1: g = dy*act'(x*scale+shift): (sum g, sum g*xhat).
template <int MODE>
__global__ void __launch_bounds__(256) colreduce_kernel(const float* __restrict__ x, int x_pitch, const float* __restrict__ dy, int dy_pitch,
                                                        long long M, int C, long long rows_per_block,
                                                        const float* __restrict__ mean, const float* __restrict__ invstd,
                                                        const float* __restrict__ scale, const float* __restrict__ shift,
                                                        int act, float slope, double* __restrict__ partial, const StatTail tail) {
  __shared__ double sh[2][8][33];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const long long r0 = blockIdx.x * rows_per_block;
  const long long r1 = (r0 + rows_per_block < M) ? r0 + rows_per_block : M;
  for (int cb = 0; cb < C; cb += 32) {
    const int c = cb + tx;
    double a0 = 0.0, a1 = 0.0;
    if (c < C) {
      float mu = 0.f, is = 0.f, sc = 0.f, sf = 0.f;
      if (MODE == 1) { mu = mean[c]; is = invstd[c]; sc = scale[c]; sf = shift[c]; }
      long long r = r0 + ty;
      while (r < r1) {
        float f0 = 0.f, f1 = 0.f;
#pragma unroll 4
        for (int u = 0; u < 16 && r < r1; ++u, r += 8) {
          float xv = x[(size_t)r * x_pitch + c];
          if (MODE == 0) {
            f0 += xv; f1 = fmaf(xv, xv, f1);
          } else {
            float g = dy[(size_t)r * dy_pitch + c] * act_grad(fmaf(xv, sc, sf), act, slope);
            f0 += g; f1 = fmaf(g, (xv - mu) * is, f1);
          }
        }
        a0 += (double)f0; a1 += (double)f1;
      }
    }
    sh[0][ty][tx] = a0; sh[1][ty][tx] = a1;
    __syncthreads();
    if (ty == 0 && c < C) {
      double s0 = 0.0, s1 = 0.0;
#pragma unroll
      for (int j = 0; j < 8; ++j) { s0 += sh[0][j][tx]; s1 += sh[1][j][tx]; }
      partial[(size_t)blockIdx.x * 2 * C + c] = s0;
      partial[(size_t)blockIdx.x * 2 * C + C + c] = s1;
    }
    __syncthreads();
  }
  stat_tail(tail, partial, C);
}


// 128-bit variant: a thread owns 4 consecutive channels and walks rows ty, ty + RP, ... of its block's row range with 8
// independent float4 loads in flight; fp32 partial sums over 8 rows are folded into double accumulators (same numerics as
// the scalar kernel); fixed summation order => deterministic.  Requires C % 4 == 0 and 16-byte aligned rows.
template <int MODE>
__global__ void __launch_bounds__(256) colreduce_v4_kernel(const float* __restrict__ x, int x_pitch, const float* __restrict__ dy, int dy_pitch,
                                                           long long M, int C, long long rows_per_block,
                                                           const float* __restrict__ mean, const float* __restrict__ invstd,
                                                           const float* __restrict__ scale, const float* __restrict__ shift,
                                                           int act, float slope, double* __restrict__ partial, const StatTail tail) {
  __shared__ double sh[256][9];
  const int Cv = C >> 2;
  const long long r0 = blockIdx.x * rows_per_block;
  const long long r1 = (r0 + rows_per_block < M) ? r0 + rows_per_block : M;
  for (int cg0 = 0; cg0 < Cv; cg0 += 256) {
    const int cw = (Cv - cg0 < 256) ? Cv - cg0 : 256;      // column groups in this pass
    const int RP = 256 / cw;                               // rows handled per pass by the block
    const int ty = threadIdx.x / cw, cg = threadIdx.x - ty * cw;
    const bool active = ty < RP;
    const int c = (cg0 + cg) << 2;
    double acc[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) acc[k] = 0.0;
    if (active) {
      float4 mu = make_float4(0, 0, 0, 0), is = mu, sc = mu, sf = mu;
      if (MODE == 1) {
        mu = *reinterpret_cast<const float4*>(mean + c); is = *reinterpret_cast<const float4*>(invstd + c);
        sc = *reinterpret_cast<const float4*>(scale + c); sf = *reinterpret_cast<const float4*>(shift + c);
      }
      for (long long r = r0 + ty; r < r1; r += 8ll * RP) {
        float4 xv[8], gv[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) {
          const long long rr = r + (long long)u * RP;
          if (rr < r1) {
            xv[u] = *reinterpret_cast<const float4*>(x + (size_t)rr * x_pitch + c);
            if (MODE == 1) gv[u] = *reinterpret_cast<const float4*>(dy + (size_t)rr * dy_pitch + c);
          } else {
            xv[u] = make_float4(0, 0, 0, 0);
            if (MODE == 1) gv[u] = make_float4(0, 0, 0, 0);
            if (MODE == 1) xv[u] = mu;          // xhat = 0 and g = 0 for the padding rows
          }
        }
        float f[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) f[k] = 0.f;
#pragma unroll
        for (int u = 0; u < 8; ++u) {
          if (MODE == 0) {
            f[0] += xv[u].x; f[1] += xv[u].y; f[2] += xv[u].z; f[3] += xv[u].w;
            f[4] = fmaf(xv[u].x, xv[u].x, f[4]); f[5] = fmaf(xv[u].y, xv[u].y, f[5]); f[6] = fmaf(xv[u].z, xv[u].z, f[6]); f[7] = fmaf(xv[u].w, xv[u].w, f[7]);
          } else {
            const float g0 = gv[u].x * act_grad(fmaf(xv[u].x, sc.x, sf.x), act, slope), g1 = gv[u].y * act_grad(fmaf(xv[u].y, sc.y, sf.y), act, slope);
            const float g2 = gv[u].z * act_grad(fmaf(xv[u].z, sc.z, sf.z), act, slope), g3 = gv[u].w * act_grad(fmaf(xv[u].w, sc.w, sf.w), act, slope);
            f[0] += g0; f[1] += g1; f[2] += g2; f[3] += g3;
            f[4] = fmaf(g0, (xv[u].x - mu.x) * is.x, f[4]); f[5] = fmaf(g1, (xv[u].y - mu.y) * is.y, f[5]);
            f[6] = fmaf(g2, (xv[u].z - mu.z) * is.z, f[6]); f[7] = fmaf(g3, (xv[u].w - mu.w) * is.w, f[7]);
          }
        }
#pragma unroll
        for (int k = 0; k < 8; ++k) acc[k] += (double)f[k];
      }
    }
#pragma unroll
    for (int k = 0; k < 8; ++k) sh[threadIdx.x][k] = acc[k];
    __syncthreads();
    if (threadIdx.x < cw) {          // thread cg sums the RP row lanes of its column group
      double s[8];
#pragma unroll
      for (int k = 0; k < 8; ++k) s[k] = 0.0;
      for (int j = 0; j < RP; ++j)
#pragma unroll
        for (int k = 0; k < 8; ++k) s[k] += sh[j * cw + threadIdx.x][k];
      double* dst = partial + (size_t)blockIdx.x * 2 * C + ((cg0 + threadIdx.x) << 2);
#pragma unroll
      for (int k = 0; k < 4; ++k) { dst[k] = s[k]; dst[C + k] = s[4 + k]; }
    }
    __syncthreads();
  }
  stat_tail(tail, partial, C);
}

// out[j] = sum_b partial[b][j], j < n2.  One warp per column group of 8: lanes = 32 interleaved slices of the block dimension, four
// independent partial sums per lane (the old one-thread-per-column loop was a chain of ~150 dependent loads: 16 us per call, 73 calls
// per step), then a fixed-order shuffle tree: deterministic.
__global__ void __launch_bounds__(256) colreduce_final_kernel(const double* __restrict__ partial, int nblocks, int n2, double* __restrict__ out) {
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int j = blockIdx.x * 8 + w;
  if (j >= n2) return;
  double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
  int b = lane;
  for (; b + 96 < nblocks; b += 128) {
    a0 += partial[(size_t)b * n2 + j]; a1 += partial[(size_t)(b + 32) * n2 + j];
    a2 += partial[(size_t)(b + 64) * n2 + j]; a3 += partial[(size_t)(b + 96) * n2 + j];
  }
  for (; b < nblocks; b += 32) a0 += partial[(size_t)b * n2 + j];
  double a = warp_sum((a0 + a1) + (a2 + a3));
  if (lane == 0) out[j] = a;
}

__global__ void bn_finalize_kernel(const double* __restrict__ sums, long long M, int C, const float* __restrict__ weight,
                                   const float* __restrict__ bias, float eps, float momentum, float* running_mean, float* running_var,
                                   float* mean_o, float* invstd_o, float* scale_o, float* shift_o) {
  int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  double m = sums[c] / (double)M;
  double var = sums[C + c] / (double)M - m * m;
  if (var < 0.0) var = 0.0;
  float is = (float)(1.0 / sqrt(var + (double)eps));
  float w = weight ? weight[c] : 1.f, b = bias ? bias[c] : 0.f;
  float sc = w * is;
  mean_o[c] = (float)m; invstd_o[c] = is; scale_o[c] = sc; shift_o[c] = b - (float)m * sc;
  if (running_mean) running_mean[c] = (1.f - momentum) * running_mean[c] + momentum * (float)m;
  if (running_var) {
    double unbiased = M > 1 ? var * ((double)M / (double)(M - 1)) : var;
    running_var[c] = (1.f - momentum) * running_var[c] + momentum * (float)unbiased;
  }
}

__global__ void bn_eval_coeffs_kernel(const float* weight, const float* bias, const float* rm, const float* rv, float eps, int C,
                                      float* scale, float* shift) {
  int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  float sc = (weight ? weight[c] : 1.f) / sqrtf(rv[c] + eps);
  scale[c] = sc;
  shift[c] = (bias ? bias[c] : 0.f) - rm[c] * sc;
}

// ---- generic NHWC elementwise drivers: VEC = 4 when C, pitches and offsets allow 128-bit access, else 1
enum { EW_AFFINE = 0, EW_BN_BWD = 1, EW_ACT_BWD = 2, EW_AXPY = 3 };
struct EwP {
  const float* a; int a_pitch;      // x / dy / dy / x
  const float* b; int b_pitch;      // - / x  / y  / -
  float* o; int o_pitch; int accumulate;
  long long M; int C;
  const float* mean; const float* invstd; const float* weight; const float* scale; const float* shift;
  const double* sums; float alpha; int act; float slope;
  __nv_bfloat16* o16; int o16_pitch;      // EW_BN_BWD, 128-bit path: the result as the bf16 tensor-core operand of the preceding convolution's gradient GEMMs (o may be NULL)
};

template <int OP>
__device__ __forceinline__ float ew_apply(const EwP& p, float a, float b, int c, float invM) {
  if (OP == EW_AFFINE) return apply_act(fmaf(a, __ldg(p.scale + c), __ldg(p.shift + c)), p.act, p.slope);
  if (OP == EW_ACT_BWD) return a * act_grad(b, p.act, p.slope);
  if (OP == EW_AXPY) return a * p.alpha;
  // EW_BN_BWD: a = dy, b = x.  dx = w*is*(g - sg - xhat*sgx) with g = dy*act'(x*sc+sf), xhat = (x-mu)*is, folded per channel into
  // dx = k1*g + k2*x + k3; the block's coefficient table [5][C] = sc | sf | k1 | k2 | k3 lives in shared memory (one 128-bit read per
  // coefficient and channel quad instead of five global and two shared scalar loads per element: the kernel was load-issue bound)
  extern __shared__ __align__(16) float ew_coef[];
  const float g = a * act_grad(fmaf(b, ew_coef[c], ew_coef[p.C + c]), p.act, p.slope);
  return fmaf(ew_coef[2 * p.C + c], g, fmaf(ew_coef[3 * p.C + c], b, ew_coef[4 * p.C + c]));
}

template <int OP, int VEC>
__global__ void __launch_bounds__(256) ew_kernel(const EwP p) {
  const int Cv = p.C / VEC;
  const long long total = p.M * Cv;
  const float invM = 1.f / (float)p.M;
  if (OP == EW_BN_BWD) {
    extern __shared__ __align__(16) float ew_coef[];
    const double inv = 1.0 / (double)p.M;
    for (int c = threadIdx.x; c < p.C; c += blockDim.x) {
      const float sg = (float)(p.sums[c] * inv), sgx = (float)(p.sums[p.C + c] * inv);
      const float is = p.invstd[c], mu = p.mean[c], w = p.weight ? p.weight[c] : 1.f;
      const float k1 = w * is, k2 = -k1 * sgx * is;
      ew_coef[c] = p.scale[c]; ew_coef[p.C + c] = p.shift[c];
      ew_coef[2 * p.C + c] = k1; ew_coef[3 * p.C + c] = k2; ew_coef[4 * p.C + c] = -k1 * sg - k2 * mu;
    }
    __syncthreads();
  }
  // 32-bit index arithmetic when the element count allows it (64-bit division costs more than the memory access it addresses)
  const bool small = total < (1ll << 31);
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
    long long m;
    int c;
    if (small) { const unsigned i32 = (unsigned)idx, m32 = i32 / (unsigned)Cv; m = m32; c = (int)(i32 - m32 * (unsigned)Cv) * VEC; }
    else { m = idx / Cv; c = (int)(idx - m * Cv) * VEC; }
    if (VEC == 4) {
      float4 a = *reinterpret_cast<const float4*>(p.a + (size_t)m * p.a_pitch + c);
      float4 b = make_float4(0.f, 0.f, 0.f, 0.f);
      if (OP == EW_BN_BWD || OP == EW_ACT_BWD) b = *reinterpret_cast<const float4*>(p.b + (size_t)m * p.b_pitch + c);
      float4 r;
      if (OP == EW_BN_BWD) {
        extern __shared__ __align__(16) float ew_coef[];
        const float4 sc = *reinterpret_cast<const float4*>(ew_coef + c), sf = *reinterpret_cast<const float4*>(ew_coef + p.C + c);
        const float4 k1 = *reinterpret_cast<const float4*>(ew_coef + 2 * p.C + c), k2 = *reinterpret_cast<const float4*>(ew_coef + 3 * p.C + c);
        const float4 k3 = *reinterpret_cast<const float4*>(ew_coef + 4 * p.C + c);
        r.x = fmaf(k1.x, a.x * act_grad(fmaf(b.x, sc.x, sf.x), p.act, p.slope), fmaf(k2.x, b.x, k3.x));
        r.y = fmaf(k1.y, a.y * act_grad(fmaf(b.y, sc.y, sf.y), p.act, p.slope), fmaf(k2.y, b.y, k3.y));
        r.z = fmaf(k1.z, a.z * act_grad(fmaf(b.z, sc.z, sf.z), p.act, p.slope), fmaf(k2.z, b.z, k3.z));
        r.w = fmaf(k1.w, a.w * act_grad(fmaf(b.w, sc.w, sf.w), p.act, p.slope), fmaf(k2.w, b.w, k3.w));
      } else if (OP == EW_AFFINE) {
        const float4 sc = __ldg(reinterpret_cast<const float4*>(p.scale + c)), sf = __ldg(reinterpret_cast<const float4*>(p.shift + c));
        r.x = apply_act(fmaf(a.x, sc.x, sf.x), p.act, p.slope); r.y = apply_act(fmaf(a.y, sc.y, sf.y), p.act, p.slope);
        r.z = apply_act(fmaf(a.z, sc.z, sf.z), p.act, p.slope); r.w = apply_act(fmaf(a.w, sc.w, sf.w), p.act, p.slope);
      } else {
        r.x = ew_apply<OP>(p, a.x, b.x, c, invM); r.y = ew_apply<OP>(p, a.y, b.y, c + 1, invM);
        r.z = ew_apply<OP>(p, a.z, b.z, c + 2, invM); r.w = ew_apply<OP>(p, a.w, b.w, c + 3, invM);
      }
      if (OP == EW_BN_BWD && p.o16) {
        const __nv_bfloat162 h0 = __floats2bfloat162_rn(r.x, r.y), h1 = __floats2bfloat162_rn(r.z, r.w);
        *reinterpret_cast<uint2*>(p.o16 + (size_t)m * p.o16_pitch + c) = make_uint2(*reinterpret_cast<const uint32_t*>(&h0), *reinterpret_cast<const uint32_t*>(&h1));
      }
      if (OP != EW_BN_BWD || p.o) {
        float4* op = reinterpret_cast<float4*>(p.o + (size_t)m * p.o_pitch + c);
        if (p.accumulate) { float4 o = *op; r.x += o.x; r.y += o.y; r.z += o.z; r.w += o.w; }
        *op = r;
      }
    } else {
      float a = p.a[(size_t)m * p.a_pitch + c];
      float b = (OP == EW_BN_BWD || OP == EW_ACT_BWD) ? p.b[(size_t)m * p.b_pitch + c] : 0.f;
      float r = ew_apply<OP>(p, a, b, c, invM);
      float* op = p.o + (size_t)m * p.o_pitch + c;
      if (p.accumulate) r += *op;
      *op = r;
    }
  }
}

static bool vec4_ok(const EwP& p) {
  auto al = [](const void* q) { return (reinterpret_cast<uintptr_t>(q) & 15) == 0; };
  bool ok = (p.C % 4 == 0) && (p.a_pitch % 4 == 0) && (p.o_pitch % 4 == 0) && al(p.a) && al(p.o);
  if (p.b) ok = ok && (p.b_pitch % 4 == 0) && al(p.b);
  if (p.scale) ok = ok && al(p.scale) && al(p.shift);
  return ok;
}

template <int OP>
static int ew_launch(const EwP& p, cudaStream_t st) {
  if (p.M == 0) return GDN_OK;
  const bool v4 = vec4_ok(p);
  long long total = p.M * (v4 ? p.C / 4 : p.C);
  const int cap = (OP == EW_BN_BWD ? 6 : 16) * kNumSMs;      // BN backward fills a per-block coefficient table first: fewer, longer blocks
  int blocks = (int)(cdiv(total, 256) < cap ? cdiv(total, 256) : cap);
  const size_t smem = OP == EW_BN_BWD ? (size_t)5 * p.C * sizeof(float) : 0;
  if (smem > 48 * 1024) { set_error("elementwise: BatchNorm backward supports at most 2457 channels"); return GDN_EINVAL; }
  if (v4) ew_kernel<OP, 4><<<blocks, 256, smem, st>>>(p);
  else ew_kernel<OP, 1><<<blocks, 256, smem, st>>>(p);
  GDN_CHECK_LAUNCH();
  return GDN_OK;
}

__global__ void bn_param_grads_kernel(const double* sums, int C, float* dweight, float* dbias) {
  int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  if (dbias) dbias[c] = (float)sums[c];
  if (dweight) dweight[c] = (float)sums[C + c];
}
__global__ void sums_to_float_kernel(const double* sums, float* out, int n, float scale) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = (float)(sums[i] * (double)scale);
}

__global__ void __launch_bounds__(256) scale_dev_kernel(const float* __restrict__ x, const float* __restrict__ sc, float* __restrict__ y, long long n, int accumulate) {
  const float a = __ldg(sc);
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    float v = a * x[i];
    y[i] = accumulate ? y[i] + v : v;
  }
}

template <int MODE>
static int colreduce(const float* x, int x_pitch, const float* dy, int dy_pitch, long long M, int C, const float* mean, const float* invstd,
                     const float* scale, const float* shift, int act, float slope, double* out, void* ws, cudaStream_t st, StatTail tail = StatTail{}) {
  StatPlan pl = stat_plan(M);
  double* partial = reinterpret_cast<double*>(ws);
  if (tail.counters) {            // single launch: group rows and ticket counters follow the partial rows in the workspace
    tail.group_partial = partial + (size_t)pl.blocks * 2 * C;
    tail.out = out;
  }
  auto al = [](const void* q) { return (reinterpret_cast<uintptr_t>(q) & 15) == 0; };
  const bool v4 = C % 4 == 0 && x_pitch % 4 == 0 && al(x) && (MODE == 0 || (dy_pitch % 4 == 0 && al(dy) && al(mean) && al(invstd) && al(scale) && al(shift)));
  if (v4) colreduce_v4_kernel<MODE><<<pl.blocks, 256, 0, st>>>(x, x_pitch, dy, dy_pitch, M, C, pl.rows_per_block, mean, invstd, scale, shift, act, slope, partial, tail);
  else colreduce_kernel<MODE><<<pl.blocks, 256, 0, st>>>(x, x_pitch, dy, dy_pitch, M, C, pl.rows_per_block, mean, invstd, scale, shift, act, slope, partial, tail);
  GDN_CHECK_LAUNCH();
  if (tail.counters) return GDN_OK;
  colreduce_final_kernel<<<(unsigned)cdiv(2 * C, 8), 256, 0, st>>>(partial, pl.blocks, 2 * C, out);
  GDN_CHECK_LAUNCH();
  return GDN_OK;
}
}  // namespace gdn

using namespace gdn;

extern "C" size_t gdn_colstats_ws_bytes(long long M, int C) {
  StatPlan pl = stat_plan(M > 0 ? M : 1);
  return (size_t)pl.blocks * 2 * (size_t)C * sizeof(double);
}
extern "C" int gdn_colstats(const float* x, int pitch, int c0, long long M, int C, double* out, void* ws, gdn_stream_t s) {
  GDN_CHECK_ARG(x && out && ws && M > 0 && C > 0 && pitch >= c0 + C);
  return colreduce<0>(x + c0, pitch, nullptr, 0, M, C, nullptr, nullptr, nullptr, nullptr, 0, 0.f, out, ws, as_stream(s));
}
extern "C" int gdn_bn_finalize(const double* sums, long long M, int C, const float* weight, const float* bias, float eps, float momentum,
                               float* running_mean, float* running_var, float* mean, float* invstd, float* scale, float* shift, gdn_stream_t s) {
  GDN_CHECK_ARG(sums && M > 0 && C > 0 && mean && invstd && scale && shift);
  bn_finalize_kernel<<<(unsigned)cdiv(C, 128), 128, 0, as_stream(s)>>>(sums, M, C, weight, bias, eps, momentum, running_mean, running_var, mean, invstd, scale, shift);
  GDN_CHECK_LAUNCH();
  return GDN_OK;
}
extern "C" size_t gdn_stat_fused_ws_bytes(long long M, int C) {
  StatPlan pl = stat_plan(M > 0 ? M : 1);
  return ((size_t)pl.blocks + (size_t)cdiv(pl.blocks, kStatGroup)) * 2 * (size_t)C * sizeof(double);
}
extern "C" int gdn_stat_fused_counters(void) { return 1 + (int)cdiv(kMaxStatBlocks, kStatGroup); }
extern "C" int gdn_bn_stats(const float* x, int pitch, int c0, long long M, int C, const float* weight, const float* bias, float eps, float momentum,
                            float* running_mean, float* running_var, long long* num_batches_tracked, float* mean, float* invstd, float* scale,
                            float* shift, void* ws, unsigned* counters, gdn_stream_t s) {
  GDN_CHECK_ARG(x && ws && counters && M > 0 && C > 0 && pitch >= c0 + C && mean && invstd && scale && shift);
  StatTail t = {};
  t.counters = counters; t.bn = 1; t.M = M; t.weight = weight; t.bias = bias; t.eps = eps; t.momentum = momentum;
  t.running_mean = running_mean; t.running_var = running_var; t.num_batches_tracked = num_batches_tracked;
  t.mean = mean; t.invstd = invstd; t.scale = scale; t.shift = shift;
  return colreduce<0>(x + c0, pitch, nullptr, 0, M, C, nullptr, nullptr, nullptr, nullptr, 0, 0.f, nullptr, ws, as_stream(s), t);
}
extern "C" int gdn_colsums_f(const float* x, int pitch, int c0, long long M, int C, double* sums, float* fsums, void* ws, unsigned* counters, gdn_stream_t s) {
  GDN_CHECK_ARG(x && ws && counters && M > 0 && C > 0 && pitch >= c0 + C && (sums || fsums));
  StatTail t = {};
  t.counters = counters; t.fout = fsums;
  return colreduce<0>(x + c0, pitch, nullptr, 0, M, C, nullptr, nullptr, nullptr, nullptr, 0, 0.f, sums, ws, as_stream(s), t);
}
extern "C" int gdn_bn_bwd_reduce_f(const float* dy, int dy_pitch, int dy_c0, const float* x, int x_pitch, int x_c0, long long M, int C,
                                   const float* mean, const float* invstd, const float* scale, const float* shift, int act, float slope,
                                   double* sums, float* fsums, void* ws, unsigned* counters, gdn_stream_t s) {
  GDN_CHECK_ARG(dy && x && mean && invstd && scale && shift && sums && ws && counters && M > 0 && C > 0);
  GDN_CHECK_ARG(dy_pitch >= dy_c0 + C && x_pitch >= x_c0 + C);
  StatTail t = {};
  t.counters = counters; t.fout = fsums;
  return colreduce<1>(x + x_c0, x_pitch, dy + dy_c0, dy_pitch, M, C, mean, invstd, scale, shift, act, slope, sums, ws, as_stream(s), t);
}
extern "C" int gdn_bn_eval_coeffs(const float* weight, const float* bias, const float* running_mean, const float* running_var, float eps,
                                  int C, float* scale, float* shift, gdn_stream_t s) {
  GDN_CHECK_ARG(running_mean && running_var && scale && shift && C > 0);
  bn_eval_coeffs_kernel<<<(unsigned)cdiv(C, 128), 128, 0, as_stream(s)>>>(weight, bias, running_mean, running_var, eps, C, scale, shift);
  GDN_CHECK_LAUNCH();
  return GDN_OK;
}
extern "C" int gdn_affine_act(const float* x, int x_pitch, int x_c0, float* y, int y_pitch, int y_c0, long long M, int C,
                              const float* scale, const float* shift, int act, float slope, gdn_stream_t s) {
  GDN_CHECK_ARG(x && y && scale && shift && M >= 0 && C > 0 && x_pitch >= x_c0 + C && y_pitch >= y_c0 + C);
  EwP p = {};
  p.a = x + x_c0; p.a_pitch = x_pitch; p.o = y + y_c0; p.o_pitch = y_pitch; p.M = M; p.C = C;
  p.scale = scale; p.shift = shift; p.act = act; p.slope = slope;
  return ew_launch<EW_AFFINE>(p, as_stream(s));
}
extern "C" int gdn_bn_bwd_reduce(const float* dy, int dy_pitch, int dy_c0, const float* x, int x_pitch, int x_c0, long long M, int C,
                                 const float* mean, const float* invstd, const float* scale, const float* shift, int act, float slope,
                                 double* sums, void* ws, gdn_stream_t s) {
  GDN_CHECK_ARG(dy && x && mean && invstd && scale && shift && sums && ws && M > 0 && C > 0);
  GDN_CHECK_ARG(dy_pitch >= dy_c0 + C && x_pitch >= x_c0 + C);
  return colreduce<1>(x + x_c0, x_pitch, dy + dy_c0, dy_pitch, M, C, mean, invstd, scale, shift, act, slope, sums, ws, as_stream(s));
}
extern "C" int gdn_bn_bwd_apply(const float* dy, int dy_pitch, int dy_c0, const float* x, int x_pitch, int x_c0,
                                float* dx, int dx_pitch, int dx_c0, int accumulate, long long M, int C,
                                const float* mean, const float* invstd, const float* weight, const float* scale, const float* shift,
                                int act, float slope, const double* sums, float* dweight, float* dbias, gdn_stream_t s) {
  GDN_CHECK_ARG(dy && x && dx && mean && invstd && scale && shift && sums && M > 0 && C > 0);
  GDN_CHECK_ARG(dy_pitch >= dy_c0 + C && x_pitch >= x_c0 + C && dx_pitch >= dx_c0 + C);
  EwP p = {};
  p.a = dy + dy_c0; p.a_pitch = dy_pitch; p.b = x + x_c0; p.b_pitch = x_pitch; p.o = dx + dx_c0; p.o_pitch = dx_pitch;
  p.accumulate = accumulate; p.M = M; p.C = C; p.mean = mean; p.invstd = invstd; p.weight = weight; p.scale = scale; p.shift = shift;
  p.sums = sums; p.act = act; p.slope = slope;
  int rc = ew_launch<EW_BN_BWD>(p, as_stream(s));
  if (rc != GDN_OK) return rc;
  if (dweight || dbias) {
    bn_param_grads_kernel<<<(unsigned)cdiv(C, 128), 128, 0, as_stream(s)>>>(sums, C, dweight, dbias);
    GDN_CHECK_LAUNCH();
  }
  return GDN_OK;
}
// The same stage 2 with the result written ONLY as bf16 [M][dx16_pitch] (round to nearest): dx of a BatchNorm that follows a bias-free
// tensor-core convolution is consumed by that convolution's gradient GEMMs alone, whose operand it is -- the fp32 tensor (4 B written by
// this kernel + 4 B read by the packing pass per element) never exists.
extern "C" int gdn_bn_bwd_apply16(const float* dy, int dy_pitch, int dy_c0, const float* x, int x_pitch, int x_c0, uint16_t* dx16, int dx16_pitch,
                                  long long M, int C, const float* mean, const float* invstd, const float* weight, const float* scale, const float* shift,
                                  int act, float slope, const double* sums, gdn_stream_t s) {
  GDN_CHECK_ARG(dy && x && dx16 && mean && invstd && scale && shift && sums && M > 0 && C > 0 && C % 8 == 0);
  GDN_CHECK_ARG(dy_pitch >= dy_c0 + C && x_pitch >= x_c0 + C && dx16_pitch >= C && dx16_pitch % 4 == 0 && ((uintptr_t)dx16 & 7) == 0);
  EwP p = {};
  p.a = dy + dy_c0; p.a_pitch = dy_pitch; p.b = x + x_c0; p.b_pitch = x_pitch; p.o = nullptr; p.o_pitch = 4;
  p.o16 = reinterpret_cast<__nv_bfloat16*>(dx16); p.o16_pitch = dx16_pitch;
  p.M = M; p.C = C; p.mean = mean; p.invstd = invstd; p.weight = weight; p.scale = scale; p.shift = shift;
  p.sums = sums; p.act = act; p.slope = slope;
  if (!vec4_ok(p)) { set_error("gdn_bn_bwd_apply16: dy / x need 16-byte aligned rows (pitches and channel offsets multiples of 4)"); return GDN_EINVAL; }
  return ew_launch<EW_BN_BWD>(p, as_stream(s));
}
extern "C" int gdn_act_bwd(const float* dy, int dy_pitch, int dy_c0, const float* y, int y_pitch, int y_c0,
                           float* dz, int dz_pitch, int dz_c0, long long M, int C, int act, float slope, gdn_stream_t s) {
  GDN_CHECK_ARG(dy && y && dz && M >= 0 && C > 0 && dy_pitch >= dy_c0 + C && y_pitch >= y_c0 + C && dz_pitch >= dz_c0 + C);
  EwP p = {};
  p.a = dy + dy_c0; p.a_pitch = dy_pitch; p.b = y + y_c0; p.b_pitch = y_pitch; p.o = dz + dz_c0; p.o_pitch = dz_pitch;
  p.M = M; p.C = C; p.act = act; p.slope = slope;
  return ew_launch<EW_ACT_BWD>(p, as_stream(s));
}
extern "C" int gdn_axpy(const float* x, int x_pitch, int x_c0, float* y, int y_pitch, int y_c0, long long M, int C, float alpha, int accumulate, gdn_stream_t s) {
  GDN_CHECK_ARG(x && y && M >= 0 && C > 0 && x_pitch >= x_c0 + C && y_pitch >= y_c0 + C);
  EwP p = {};
  p.a = x + x_c0; p.a_pitch = x_pitch; p.o = y + y_c0; p.o_pitch = y_pitch; p.accumulate = accumulate; p.M = M; p.C = C; p.alpha = alpha;
  return ew_launch<EW_AXPY>(p, as_stream(s));
}
extern "C" int gdn_sums_to_float(const double* sums, float* out, int n, float scale, gdn_stream_t s) {
  GDN_CHECK_ARG(sums && out && n > 0);
  sums_to_float_kernel<<<(unsigned)cdiv(n, 128), 128, 0, as_stream(s)>>>(sums, out, n, scale);
  GDN_CHECK_LAUNCH();
  return GDN_OK;
}
extern "C" int gdn_scale_dev(const float* x, const float* scalar, float* y, long long n, int accumulate, gdn_stream_t s) {
  GDN_CHECK_ARG(x && scalar && y && n >= 0);
  if (n == 0) return GDN_OK;
  int blocks = (int)(cdiv(n, 256) < 16 * kNumSMs ? cdiv(n, 256) : 16 * kNumSMs);
  scale_dev_kernel<<<blocks, 256, 0, as_stream(s)>>>(x, scalar, y, n, accumulate);
  GDN_CHECK_LAUNCH();
  return GDN_OK;
}
