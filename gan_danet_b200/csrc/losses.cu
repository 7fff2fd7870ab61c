// Scalar losses (+ their gradients in the same pass), dot products and the fused AdamW update.
// Reference call sites: nn.MSELoss / nn.BCEWithLogitsLoss at /root/reference/GAN_DANet_train.ipynb:190-191,252-253,261-262;
// TVLoss models/losses.py:76-87; SSIM models/losses.py:90-147; F.l1_loss models/losses.py:72;
// torch.optim.AdamW GAN_DANet_train.ipynb:182-183,256,269.
// All reductions: per-block partial sums in double -> one finishing block (fixed order => bitwise deterministic).
#include <math.h>
#include "common.cuh"

namespace gdn {

constexpr int kRedBlocks = 8 * kNumSMs;   // upper bound on stage-1 blocks
constexpr int kRedSlots = 4;              // doubles per block

static inline int red_blocks(long long n, int per_thread = 8) {
  long long b = cdiv(n, 256ll * per_thread);
  if (b < 1) b = 1;
  return (int)(b < kRedBlocks ? b : kRedBlocks);
}

__device__ __forceinline__ void block_store_partials(double v0, double v1, double* partial) {
  __shared__ double sh[32];
  double s0 = block_sum<double>(v0, sh);
  double s1 = block_sum<double>(v1, sh);
  if (threadIdx.x == 0) { partial[(size_t)blockIdx.x * kRedSlots] = s0; partial[(size_t)blockIdx.x * kRedSlots + 1] = s1; }
}

// loss[0] (+)= c0*sum(slot0) + c1*sum(slot1)
__global__ void __launch_bounds__(256) finish_kernel(const double* __restrict__ partial, int nblocks, double c0, double c1, float* loss, int accumulate) {
  __shared__ double sh[32];
  double a0 = 0.0, a1 = 0.0;
  for (int b = threadIdx.x; b < nblocks; b += blockDim.x) { a0 += partial[(size_t)b * kRedSlots]; a1 += partial[(size_t)b * kRedSlots + 1]; }
  a0 = block_sum<double>(a0, sh);
  a1 = block_sum<double>(a1, sh);
  if (threadIdx.x == 0) {
    float v = (float)(c0 * a0 + c1 * a1);
    loss[0] = accumulate ? loss[0] + v : v;
  }
}

// MODE 0: MSE, MODE 1: L1
template <int MODE>
__global__ void __launch_bounds__(256) pair_loss_kernel(const float* __restrict__ a, const float* __restrict__ b, long long n, float* __restrict__ grad,
                                                        float gcoef, int accumulate, double* __restrict__ partial) {
  float acc = 0.f; double dacc = 0.0; int cnt = 0;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    float d = a[i] - b[i];
    float g;
    if (MODE == 0) { acc = fmaf(d, d, acc); g = d * gcoef; }
    else { acc += fabsf(d); g = (d > 0.f ? gcoef : (d < 0.f ? -gcoef : 0.f)); }
    if (grad) grad[i] = accumulate ? grad[i] + g : g;
    if (++cnt == 32) { dacc += (double)acc; acc = 0.f; cnt = 0; }
  }
  dacc += (double)acc;
  block_store_partials(dacc, 0.0, partial);
}

// 128-bit variant (n % 4 == 0, 16-byte aligned pointers): the scalar kernel is instruction bound on the 1 GB VGG feature maps (4.2 TB/s)
template <int MODE>
__global__ void __launch_bounds__(256) pair_loss_v4_kernel(const float* __restrict__ a, const float* __restrict__ b, long long n4, float* __restrict__ grad,
                                                           float gcoef, int accumulate, double* __restrict__ partial) {
  float acc = 0.f; double dacc = 0.0; int cnt = 0;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
    const float4 x = __ldg(reinterpret_cast<const float4*>(a) + i), y = __ldg(reinterpret_cast<const float4*>(b) + i);
    const float d[4] = {x.x - y.x, x.y - y.y, x.z - y.z, x.w - y.w};
    float g[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      if (MODE == 0) { acc = fmaf(d[e], d[e], acc); g[e] = d[e] * gcoef; }
      else { acc += fabsf(d[e]); g[e] = (d[e] > 0.f ? gcoef : (d[e] < 0.f ? -gcoef : 0.f)); }
    }
    if (grad) {
      float4* gp = reinterpret_cast<float4*>(grad) + i;
      if (accumulate) { const float4 o = *gp; g[0] += o.x; g[1] += o.y; g[2] += o.z; g[3] += o.w; }
      *gp = make_float4(g[0], g[1], g[2], g[3]);
    }
    if (++cnt == 8) { dacc += (double)acc; acc = 0.f; cnt = 0; }
  }
  dacc += (double)acc;
  block_store_partials(dacc, 0.0, partial);
}

__global__ void __launch_bounds__(256) tv_kernel(const float* __restrict__ x, int B, int H, int W, float* __restrict__ grad, float ch, float cw,
                                                  int accumulate, double* __restrict__ partial) {
  const long long n = (long long)B * H * W;
  double dh = 0.0, dw = 0.0; float fh = 0.f, fw = 0.f; int cnt = 0;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    int w = (int)(i % W); int h = (int)((i / W) % H);
    float c = x[i];
    float up = h > 0 ? c - x[i - W] : 0.f;        // x[h]-x[h-1]
    float dn = h < H - 1 ? x[i + W] - c : 0.f;    // x[h+1]-x[h]
    float lf = w > 0 ? c - x[i - 1] : 0.f;
    float rt = w < W - 1 ? x[i + 1] - c : 0.f;
    fh = fmaf(dn, dn, fh); fw = fmaf(rt, rt, fw);
    if (grad) { float g = ch * (up - dn) + cw * (lf - rt); grad[i] = accumulate ? grad[i] + g : g; }
    if (++cnt == 32) { dh += (double)fh; dw += (double)fw; fh = fw = 0.f; cnt = 0; }
  }
  dh += (double)fh; dw += (double)fw;
  block_store_partials(dh, dw, partial);
}

__global__ void __launch_bounds__(256) bce_kernel(const float* __restrict__ z, int n, const float* __restrict__ target_ptr, float target_const, float* loss, float* grad, float gscale) {
  __shared__ double sh[32];
  double acc = 0.0;
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    float v = z[i];
    const float target = target_ptr ? target_ptr[i] : target_const;
    acc += (double)(fmaxf(v, 0.f) - v * target + log1pf(expf(-fabsf(v))));
    if (grad) grad[i] = gscale * (1.f / (1.f + expf(-v)) - target) / (float)n;
  }
  acc = block_sum<double>(acc, sh);
  if (threadIdx.x == 0) loss[0] = (float)(acc / (double)n);
}

__global__ void __launch_bounds__(256) dot_kernel(const float* __restrict__ a, int a_pitch, const float* __restrict__ b, int b_pitch, long long M, int C,
                                                   double* __restrict__ partial) {
  const long long n = M * C;
  float acc = 0.f; double dacc = 0.0; int cnt = 0;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    long long m = i / C; int c = (int)(i - m * C);
    acc = fmaf(a[(size_t)m * a_pitch + c], b[(size_t)m * b_pitch + c], acc);
    if (++cnt == 32) { dacc += (double)acc; acc = 0.f; cnt = 0; }
  }
  dacc += (double)acc;
  block_store_partials(dacc, 0.0, partial);
}

// ---- SSIM forward: 32x16 output tile, 11x11 separable-by-outer-product Gaussian taps from shared memory
struct SsimWin { float g[11]; };
constexpr int ST_W = 32, ST_H = 16, SHALO = 5;
__global__ void __launch_bounds__(ST_W* ST_H) ssim_kernel(const float* __restrict__ a, const float* __restrict__ b, int B, int H, int W, SsimWin win,
                                                           double* __restrict__ partial) {
  __shared__ float sa[ST_H + 2 * SHALO][ST_W + 2 * SHALO + 1];
  __shared__ float sb[ST_H + 2 * SHALO][ST_W + 2 * SHALO + 1];
  const int tilesx = (W + ST_W - 1) / ST_W, tilesy = (H + ST_H - 1) / ST_H;
  const long long ntiles = (long long)B * tilesx * tilesy;
  const int tx = threadIdx.x % ST_W, ty = threadIdx.x / ST_W;
  double dacc = 0.0;
  for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    int bx = (int)(tile % tilesx); long long r = tile / tilesx; int by = (int)(r % tilesy); int bb = (int)(r / tilesy);
    const float* pa = a + (size_t)bb * H * W; const float* pb = b + (size_t)bb * H * W;
    __syncthreads();
    for (int i = threadIdx.x; i < (ST_H + 2 * SHALO) * (ST_W + 2 * SHALO); i += blockDim.x) {
      int ly = i / (ST_W + 2 * SHALO), lx = i % (ST_W + 2 * SHALO);
      int gy = by * ST_H + ly - SHALO, gx = bx * ST_W + lx - SHALO;
      bool ok = gy >= 0 && gy < H && gx >= 0 && gx < W;   // zero padding (F.conv2d padding=5)
      sa[ly][lx] = ok ? pa[(size_t)gy * W + gx] : 0.f;
      sb[ly][lx] = ok ? pb[(size_t)gy * W + gx] : 0.f;
    }
    __syncthreads();
    int oy = by * ST_H + ty, ox = bx * ST_W + tx;
    if (oy < H && ox < W) {
      float m1 = 0.f, m2 = 0.f, s11 = 0.f, s22 = 0.f, s12 = 0.f;
#pragma unroll
      for (int i = 0; i < 11; ++i) {
#pragma unroll
        for (int j = 0; j < 11; ++j) {
          float wgt = win.g[i] * win.g[j];
          float u = sa[ty + i][tx + j], v = sb[ty + i][tx + j];
          m1 = fmaf(wgt, u, m1); m2 = fmaf(wgt, v, m2);
          s11 = fmaf(wgt, u * u, s11); s22 = fmaf(wgt, v * v, s22); s12 = fmaf(wgt, u * v, s12);
        }
      }
      float mu1sq = m1 * m1, mu2sq = m2 * m2, mu12 = m1 * m2;
      float sig1 = s11 - mu1sq, sig2 = s22 - mu2sq, sig12 = s12 - mu12;
      const float c1 = 0.01f * 0.01f, c2 = 0.03f * 0.03f;
      float v = ((2.f * mu12 + c1) * (2.f * sig12 + c2)) / ((mu1sq + mu2sq + c1) * (sig1 + sig2 + c2));
      dacc += (double)v;
    }
  }
  block_store_partials(dacc, 0.0, partial);
}

// ---- AdamW (decoupled weight decay), torch.optim.AdamW single-tensor formula
// dyn != nullptr: lr, step_size (= lr / (1 - beta1^step)) and bc2_sqrt (= sqrt(1 - beta2^step)) are read from that device array instead of the
// by-value fields -- the step-dependent scalars of a CUDA-graph-captured optimiser step (the graph bakes kernel arguments, not device memory)
struct AdamP { float lr, beta1, beta2, eps, wd, step_size, bc2_sqrt, grad_scale; const float* dyn; };
__device__ __forceinline__ void adam_dyn(AdamP& a) {
  if (a.dyn) { a.lr = __ldg(a.dyn); a.step_size = __ldg(a.dyn + 1); a.bc2_sqrt = __ldg(a.dyn + 2); }
}
template <int VEC>
__global__ void __launch_bounds__(256) adamw_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m, float* __restrict__ v, long long n, AdamP a) {
  adam_dyn(a);
  const long long nv = n / VEC;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < nv; i += (long long)gridDim.x * blockDim.x) {
    float pv[VEC], gv[VEC], mv[VEC], vv[VEC];
    if (VEC == 4) {
      *reinterpret_cast<float4*>(pv) = reinterpret_cast<const float4*>(p)[i];
      *reinterpret_cast<float4*>(gv) = reinterpret_cast<const float4*>(g)[i];
      *reinterpret_cast<float4*>(mv) = reinterpret_cast<const float4*>(m)[i];
      *reinterpret_cast<float4*>(vv) = reinterpret_cast<const float4*>(v)[i];
    } else { pv[0] = p[i]; gv[0] = g[i]; mv[0] = m[i]; vv[0] = v[i]; }
#pragma unroll
    for (int k = 0; k < VEC; ++k) {
      float gr = gv[k] * a.grad_scale;
      float pp = pv[k] * (1.f - a.lr * a.wd);
      float mm = mv[k] + (gr - mv[k]) * (1.f - a.beta1);
      float v2 = vv[k] * a.beta2 + (1.f - a.beta2) * gr * gr;
      float denom = sqrtf(v2) / a.bc2_sqrt + a.eps;
      pv[k] = pp - a.step_size * (mm / denom);
      mv[k] = mm; vv[k] = v2;
    }
    if (VEC == 4) {
      reinterpret_cast<float4*>(p)[i] = *reinterpret_cast<float4*>(pv);
      reinterpret_cast<float4*>(m)[i] = *reinterpret_cast<float4*>(mv);
      reinterpret_cast<float4*>(v)[i] = *reinterpret_cast<float4*>(vv);
    } else { p[i] = pv[0]; m[i] = mv[0]; v[i] = vv[0]; }
  }
}
// Many small tensors in one launch (the generator has ~100 parameter tensors of 24 .. 500 K elements: one launch each is pure launch
// latency).  blockIdx.y = tensor, blockIdx.x strides over it.  Same arithmetic as adamw_kernel, scalar accesses (no alignment demands).
constexpr int kAdamMulti = 64;
struct AdamMultiP { float* p[kAdamMulti]; const float* g[kAdamMulti]; float* m[kAdamMulti]; float* v[kAdamMulti]; int n[kAdamMulti]; };
__global__ void __launch_bounds__(256) adamw_multi_kernel(const __grid_constant__ AdamMultiP t, AdamP a) {
  adam_dyn(a);
  const int k = blockIdx.y;
  float* __restrict__ p = t.p[k]; const float* __restrict__ g = t.g[k]; float* __restrict__ m = t.m[k]; float* __restrict__ v = t.v[k];
  const int n = t.n[k];
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const float gr = g[i] * a.grad_scale;
    const float pp = p[i] * (1.f - a.lr * a.wd);
    const float mm = m[i] + (gr - m[i]) * (1.f - a.beta1);
    const float v2 = v[i] * a.beta2 + (1.f - a.beta2) * gr * gr;
    const float denom = sqrtf(v2) / a.bc2_sqrt + a.eps;
    p[i] = pp - a.step_size * (mm / denom);
    m[i] = mm; v[i] = v2;
  }
}
}  // namespace gdn

using namespace gdn;

extern "C" size_t gdn_dot_ws_bytes(long long n) { (void)n; return (size_t)kRedBlocks * kRedSlots * sizeof(double); }

// 128-bit variant (C, pitches and offsets multiples of 4, 16-byte aligned bases, M*C/4 < 2^31): two independent float4 pairs in flight per thread
namespace gdn {
__global__ void __launch_bounds__(256) dot_v4_kernel(const float* __restrict__ a, int a_pitch, const float* __restrict__ b, int b_pitch, int M, int Cv,
                                                      double* __restrict__ partial) {
  const int n = M * Cv;
  const int stride = gridDim.x * blockDim.x;
  float acc0 = 0.f, acc1 = 0.f; double dacc = 0.0; int cnt = 0;
  auto term = [&](int i) {
    const int m = i / Cv, c = (i - m * Cv) << 2;
    const float4 x = __ldg(reinterpret_cast<const float4*>(a + (size_t)m * a_pitch + c)), y = __ldg(reinterpret_cast<const float4*>(b + (size_t)m * b_pitch + c));
    return fmaf(x.x, y.x, fmaf(x.y, y.y, fmaf(x.z, y.z, x.w * y.w)));
  };
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  for (; i + stride < n; i += 2 * stride) {
    const float t0 = term(i), t1 = term(i + stride);
    acc0 += t0; acc1 += t1;
    if (++cnt == 8) { dacc += (double)acc0 + (double)acc1; acc0 = acc1 = 0.f; cnt = 0; }
  }
  if (i < n) acc0 += term(i);
  dacc += (double)acc0 + (double)acc1;
  block_store_partials(dacc, 0.0, partial);
}
}  // namespace gdn

extern "C" int gdn_dot(const float* a, int a_pitch, int a_c0, const float* b, int b_pitch, int b_c0, long long M, int C, float* out, void* ws, gdn_stream_t s) {
  GDN_CHECK_ARG(a && b && out && ws && M > 0 && C > 0 && a_pitch >= a_c0 + C && b_pitch >= b_c0 + C);
  int blocks = red_blocks(M * C);
  double* partial = reinterpret_cast<double*>(ws);
  const bool v4 = C % 4 == 0 && a_pitch % 4 == 0 && b_pitch % 4 == 0 && a_c0 % 4 == 0 && b_c0 % 4 == 0 && ((uintptr_t)a & 15) == 0 && ((uintptr_t)b & 15) == 0 &&
                  M * (C / 4) < (1ll << 30);
  if (v4) dot_v4_kernel<<<blocks, 256, 0, as_stream(s)>>>(a + a_c0, a_pitch, b + b_c0, b_pitch, (int)M, C / 4, partial);
  else dot_kernel<<<blocks, 256, 0, as_stream(s)>>>(a + a_c0, a_pitch, b + b_c0, b_pitch, M, C, partial);
  GDN_CHECK_LAUNCH();
  finish_kernel<<<1, 256, 0, as_stream(s)>>>(partial, blocks, 1.0, 0.0, out, 0);
  GDN_CHECK_LAUNCH();
  return GDN_OK;
}

static inline bool al16p(const void* q) { return (reinterpret_cast<uintptr_t>(q) & 15) == 0; }

extern "C" int gdn_mse(const float* a, const float* b, long long n, float* loss, float* grad, float gscale, int accumulate, void* ws, gdn_stream_t s) {
  GDN_CHECK_ARG(a && b && loss && ws && n > 0);
  int blocks = red_blocks(n);
  double* partial = reinterpret_cast<double*>(ws);
  if (n % 4 == 0 && al16p(a) && al16p(b) && al16p(grad)) pair_loss_v4_kernel<0><<<blocks, 256, 0, as_stream(s)>>>(a, b, n / 4, grad, gscale * 2.f / (float)n, accumulate, partial);
  else pair_loss_kernel<0><<<blocks, 256, 0, as_stream(s)>>>(a, b, n, grad, gscale * 2.f / (float)n, accumulate, partial);
  GDN_CHECK_LAUNCH();
  finish_kernel<<<1, 256, 0, as_stream(s)>>>(partial, blocks, 1.0 / (double)n, 0.0, loss, 0);
  GDN_CHECK_LAUNCH();
  return GDN_OK;
}
extern "C" int gdn_l1(const float* a, const float* b, long long n, float* loss, int loss_accumulate, float* grad, float gscale, int accumulate, void* ws, gdn_stream_t s) {
  GDN_CHECK_ARG(a && b && loss && ws && n > 0);
  int blocks = red_blocks(n);
  double* partial = reinterpret_cast<double*>(ws);
  if (n % 4 == 0 && al16p(a) && al16p(b) && al16p(grad)) pair_loss_v4_kernel<1><<<blocks, 256, 0, as_stream(s)>>>(a, b, n / 4, grad, gscale / (float)n, accumulate, partial);
  else pair_loss_kernel<1><<<blocks, 256, 0, as_stream(s)>>>(a, b, n, grad, gscale / (float)n, accumulate, partial);
  GDN_CHECK_LAUNCH();
  finish_kernel<<<1, 256, 0, as_stream(s)>>>(partial, blocks, 1.0 / (double)n, 0.0, loss, loss_accumulate);
  GDN_CHECK_LAUNCH();
  return GDN_OK;
}
extern "C" int gdn_tv(const float* x, int B, int H, int W, float weight, float* loss, float* grad, float gscale, int accumulate, void* ws, gdn_stream_t s) {
  GDN_CHECK_ARG(x && loss && ws && B > 0 && H > 1 && W > 1);
  long long n = (long long)B * H * W;
  int blocks = red_blocks(n);
  double* partial = reinterpret_cast<double*>(ws);
  double count_h = (double)B * (H - 1) * W, count_w = (double)B * H * (W - 1);
  double kh = (double)weight * 2.0 / count_h / (double)B, kw = (double)weight * 2.0 / count_w / (double)B;
  tv_kernel<<<blocks, 256, 0, as_stream(s)>>>(x, B, H, W, grad, (float)(2.0 * kh * gscale), (float)(2.0 * kw * gscale), accumulate, partial);
  GDN_CHECK_LAUNCH();
  finish_kernel<<<1, 256, 0, as_stream(s)>>>(partial, blocks, kh, kw, loss, 0);
  GDN_CHECK_LAUNCH();
  return GDN_OK;
}
extern "C" int gdn_bce_logits(const float* z, int n, const float* target_ptr, float target, float* loss, float* grad, float gscale, gdn_stream_t s) {
  GDN_CHECK_ARG(z && loss && n > 0);
  bce_kernel<<<1, 256, 0, as_stream(s)>>>(z, n, target_ptr, target, loss, grad, gscale);
  GDN_CHECK_LAUNCH();
  return GDN_OK;
}
extern "C" size_t gdn_ssim_ws_bytes(int B, int H, int W) { (void)B; (void)H; (void)W; return (size_t)kRedBlocks * kRedSlots * sizeof(double); }
extern "C" int gdn_ssim(const float* a, const float* b, int B, int H, int W, float* out, void* ws, gdn_stream_t s) {
  GDN_CHECK_ARG(a && b && out && ws && B > 0 && H > 0 && W > 0);
  SsimWin win;
  float sum = 0.f;
  for (int i = 0; i < 11; ++i) { win.g[i] = expf(-(float)((i - 5) * (i - 5)) / (2.f * 1.5f * 1.5f)); sum += win.g[i]; }
  for (int i = 0; i < 11; ++i) win.g[i] /= sum;
  long long ntiles = (long long)B * cdiv(W, ST_W) * cdiv(H, ST_H);
  int blocks = (int)(ntiles < kRedBlocks ? ntiles : kRedBlocks);
  double* partial = reinterpret_cast<double*>(ws);
  ssim_kernel<<<blocks, ST_W * ST_H, 0, as_stream(s)>>>(a, b, B, H, W, win, partial);
  GDN_CHECK_LAUNCH();
  finish_kernel<<<1, 256, 0, as_stream(s)>>>(partial, blocks, 1.0 / ((double)B * H * W), 0.0, out, 0);
  GDN_CHECK_LAUNCH();
  return GDN_OK;
}

static int adamw_multi_impl(int count, float* const* p, const float* const* g, float* const* m, float* const* v, const long long* n, float lr, float beta1,
                            float beta2, float eps, float wd, int step, float grad_scale, const float* dyn, gdn_stream_t s) {
  GDN_CHECK_ARG(count > 0 && p && g && m && v && n && (step >= 1 || dyn));
  AdamP a;
  a.lr = lr; a.beta1 = beta1; a.beta2 = beta2; a.eps = eps; a.wd = wd; a.grad_scale = grad_scale; a.dyn = dyn;
  if (step < 1) step = 1;
  double bc1 = 1.0 - pow((double)beta1, (double)step), bc2 = 1.0 - pow((double)beta2, (double)step);
  a.step_size = (float)((double)lr / bc1);
  a.bc2_sqrt = (float)sqrt(bc2);
  for (int i0 = 0; i0 < count; i0 += kAdamMulti) {
    AdamMultiP t;
    const int c = count - i0 < kAdamMulti ? count - i0 : kAdamMulti;
    long long nmax = 0;
    for (int i = 0; i < c; ++i) {
      GDN_CHECK_ARG(p[i0 + i] && g[i0 + i] && m[i0 + i] && v[i0 + i] && n[i0 + i] > 0 && n[i0 + i] < (1ll << 31));
      t.p[i] = p[i0 + i]; t.g[i] = g[i0 + i]; t.m[i] = m[i0 + i]; t.v[i] = v[i0 + i]; t.n[i] = (int)n[i0 + i];
      if (n[i0 + i] > nmax) nmax = n[i0 + i];
    }
    for (int i = c; i < kAdamMulti; ++i) { t.p[i] = nullptr; t.g[i] = nullptr; t.m[i] = nullptr; t.v[i] = nullptr; t.n[i] = 0; }
    long long bx = cdiv(nmax, 256 * 8);
    if (bx < 1) bx = 1;
    if (bx > 64) bx = 64;
    adamw_multi_kernel<<<dim3((unsigned)bx, (unsigned)c), 256, 0, as_stream(s)>>>(t, a);
    GDN_CHECK_LAUNCH();
  }
  return GDN_OK;
}

extern "C" int gdn_adamw_multi(int count, float* const* p, const float* const* g, float* const* m, float* const* v, const long long* n, float lr, float beta1,
                               float beta2, float eps, float wd, int step, float grad_scale, gdn_stream_t s) {
  return adamw_multi_impl(count, p, g, m, v, n, lr, beta1, beta2, eps, wd, step, grad_scale, nullptr, s);
}
extern "C" int gdn_adamw_multi_dyn(int count, float* const* p, const float* const* g, float* const* m, float* const* v, const long long* n, const float* dyn,
                                   float beta1, float beta2, float eps, float wd, float grad_scale, gdn_stream_t s) {
  GDN_CHECK_ARG(dyn != nullptr);
  return adamw_multi_impl(count, p, g, m, v, n, 0.f, beta1, beta2, eps, wd, 0, grad_scale, dyn, s);
}

static int adamw_impl(float* p, const float* g, float* m, float* v, long long n, float lr, float beta1, float beta2, float eps, float wd,
                      int step, float grad_scale, const float* dyn, gdn_stream_t s) {
  GDN_CHECK_ARG(p && g && m && v && n > 0 && (step >= 1 || dyn));
  AdamP a;
  a.lr = lr; a.beta1 = beta1; a.beta2 = beta2; a.eps = eps; a.wd = wd; a.grad_scale = grad_scale; a.dyn = dyn;
  if (step < 1) step = 1;
  double bc1 = 1.0 - pow((double)beta1, (double)step), bc2 = 1.0 - pow((double)beta2, (double)step);
  a.step_size = (float)((double)lr / bc1);
  a.bc2_sqrt = (float)sqrt(bc2);
  auto al = [](const void* q) { return (reinterpret_cast<uintptr_t>(q) & 15) == 0; };
  cudaStream_t st = as_stream(s);
  if (n % 4 == 0 && al(p) && al(g) && al(m) && al(v)) {
    long long nv = n / 4;
    int blocks = (int)(cdiv(nv, 256) < 16 * kNumSMs ? cdiv(nv, 256) : 16 * kNumSMs);
    adamw_kernel<4><<<blocks, 256, 0, st>>>(p, g, m, v, n, a);
  } else {
    int blocks = (int)(cdiv(n, 256) < 16 * kNumSMs ? cdiv(n, 256) : 16 * kNumSMs);
    adamw_kernel<1><<<blocks, 256, 0, st>>>(p, g, m, v, n, a);
  }
  GDN_CHECK_LAUNCH();
  return GDN_OK;
}

extern "C" int gdn_adamw(float* p, const float* g, float* m, float* v, long long n, float lr, float beta1, float beta2, float eps, float wd,
                         int step, float grad_scale, gdn_stream_t s) {
  return adamw_impl(p, g, m, v, n, lr, beta1, beta2, eps, wd, step, grad_scale, nullptr, s);
}
extern "C" int gdn_adamw_dyn(float* p, const float* g, float* m, float* v, long long n, const float* dyn, float beta1, float beta2, float eps, float wd,
                             float grad_scale, gdn_stream_t s) {
  GDN_CHECK_ARG(dyn != nullptr);
  return adamw_impl(p, g, m, v, n, 0.f, beta1, beta2, eps, wd, 0, grad_scale, dyn, s);
}
