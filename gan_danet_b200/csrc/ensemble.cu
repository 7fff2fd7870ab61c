// Deep-ensemble statistics and inference post-processing on the device (f1, f3).
// Reference call sites (CPU numpy in the reference):
//   deep_ensemble.ipynb:415-416  grace_scaler.inverse_transform of member predictions          -> gdn_destandardise
//   deep_ensemble.ipynb:450-463  mask == 0 -> NaN, np.nanmean over (lat, lon) per member/month  -> gdn_masked_spatial_mean
//   deep_ensemble.ipynb:466-467  np.nanmean / np.nanstd over the members                        -> gdn_ensemble_stats
//   test.ipynb:115-125,175       mild_histogram_matching(source, reference, weight)             -> gdn_hist_match
//   test.ipynb:180-191           + trend, scaler.inverse_transform, tpb_h == 0 -> NaN           -> gdn_destandardise
// All kernels are HBM-bound streaming passes: every field is read once and written at most once (4 B per element each way);
// grids are sized in multiples of the 148 SMs; accumulations run in a fixed order (bitwise reproducible).
#include <math.h>
#include "common.cuh"

using namespace gdn;

namespace {

// ------------------------------------------------------------------------------------------------ de-standardise (+ trend, + mask)
// out[b][p] = keep[p] ? (x[b][p] + trend[b][p]) * scale + shift : NaN       (trend, keep optional)
__global__ void __launch_bounds__(256) destandardise_kernel(const float* __restrict__ x, const float* __restrict__ trend, const unsigned char* __restrict__ keep,
                                                            float* __restrict__ out, long long rows, long long HW, float scale, float shift) {
  const long long total = rows * HW;
  const float nanv = __int_as_float(0x7fc00000);
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    float v = __ldg(x + i);
    if (trend) v += __ldg(trend + i);
    v = fmaf(v, scale, shift);
    if (keep && !keep[i % HW]) v = nanv;
    out[i] = v;
  }
}

// ------------------------------------------------------------------------------------------------ masked spatial mean
// out[r] = nanmean_p { x[r][p]*scale + shift : keep[p] != 0 }  (NaN inputs are skipped like np.nanmean; no valid pixel -> NaN).
// One CTA per row: 256 threads stride the H*W pixels (coalesced), per-thread partial sums in double, fixed-order block reduction.
__global__ void __launch_bounds__(256) masked_spatial_mean_kernel(const float* __restrict__ x, const unsigned char* __restrict__ keep, long long HW, float scale,
                                                                  float shift, float* __restrict__ out) {
  __shared__ double red[32];
  const float* row = x + (long long)blockIdx.x * HW;
  double s = 0.0, c = 0.0;
  for (long long p = threadIdx.x; p < HW; p += blockDim.x) {
    const float v = __ldg(row + p);
    if ((keep == nullptr || keep[p]) && v == v) { s += (double)fmaf(v, scale, shift); c += 1.0; }
  }
  s = block_sum(s, red);
  c = block_sum(c, red);
  if (threadIdx.x == 0) out[blockIdx.x] = c > 0.0 ? (float)(s / c) : __int_as_float(0x7fc00000);
}

// ------------------------------------------------------------------------------------------------ statistics over the members
// preds: M fields of n elements (member m at preds + m*stride).  mean[i], std[i] over the non-NaN members (ddof = 0, np.nanstd);
// two passes over the members held in registers for M <= 16 (the one-per-GPU ensemble), re-read otherwise.
template <int MAXM>
__global__ void __launch_bounds__(256) ensemble_stats_kernel(const float* __restrict__ preds, long long stride, int M, long long n, float scale, float shift,
                                                             float* __restrict__ mean, float* __restrict__ stdev) {
  const float nanv = __int_as_float(0x7fc00000);
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    float v[MAXM > 0 ? MAXM : 1];
    double s = 0.0;
    int c = 0;
    if (MAXM > 0) {
#pragma unroll
      for (int m = 0; m < MAXM; ++m) {
        v[m] = m < M ? fmaf(__ldg(preds + m * stride + i), scale, shift) : nanv;
        if (v[m] == v[m]) { s += (double)v[m]; ++c; }
      }
    } else {
      for (int m = 0; m < M; ++m) {
        const float t = fmaf(__ldg(preds + m * stride + i), scale, shift);
        if (t == t) { s += (double)t; ++c; }
      }
    }
    if (c == 0) { mean[i] = nanv; if (stdev) stdev[i] = nanv; continue; }
    const double mu = s / c;
    mean[i] = (float)mu;
    if (!stdev) continue;
    double q = 0.0;
    if (MAXM > 0) {
#pragma unroll
      for (int m = 0; m < MAXM; ++m)
        if (v[m] == v[m]) { const double dlt = (double)v[m] - mu; q += dlt * dlt; }
    } else {
      for (int m = 0; m < M; ++m) {
        const float t = fmaf(__ldg(preds + m * stride + i), scale, shift);
        if (t == t) { const double dlt = (double)t - mu; q += dlt * dlt; }
      }
    }
    stdev[i] = (float)sqrt(q / c);
  }
}

// ------------------------------------------------------------------------------------------------ histogram matching
// number of elements <= v / < v in an ascending array
__device__ __forceinline__ int upper_bound(const float* __restrict__ a, int n, float v) {
  int lo = 0, hi = n;
  while (lo < hi) { const int mid = (lo + hi) >> 1; if (a[mid] <= v) lo = mid + 1; else hi = mid; }
  return lo;
}
__device__ __forceinline__ int lower_bound(const float* __restrict__ a, int n, float v) {
  int lo = 0, hi = n;
  while (lo < hi) { const int mid = (lo + hi) >> 1; if (a[mid] < v) lo = mid + 1; else hi = mid; }
  return lo;
}

// One thread per source element.  With ascending copies S (source) and T (reference) of one sample:
//   s_q = #{S <= v} / n_s                                    (np.unique counts + cumsum of the source, test.ipynb:118-120)
//   the knots of the reference CDF are the LAST element of every run of equal values of T: (x_j, f_j) = ((j+1)/n_t, T[j])   (:119,121)
//   matched = np.interp(s_q, x, f)  (clamped below x_0; piecewise linear in float64)                                        (:122-123)
//   out = (1 - weight) * v + weight * matched                                                                               (:124)
__global__ void __launch_bounds__(256) hist_match_kernel(const float* __restrict__ src, const float* __restrict__ src_sorted, const float* __restrict__ ref_sorted,
                                                         float* __restrict__ out, int ns, int nt, float weight) {
  const float* S = src_sorted + (long long)blockIdx.y * ns;
  const float* T = ref_sorted + (long long)blockIdx.y * nt;
  const float* x = src + (long long)blockIdx.y * ns;
  float* o = out + (long long)blockIdx.y * ns;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < ns; i += gridDim.x * blockDim.x) {
    const float v = x[i];
    const double sq = (double)upper_bound(S, ns, v) / (double)ns;
    // largest knot j with (j+1)/nt <= sq
    int j0 = (int)floor(sq * (double)nt) - 1;
    while (j0 + 1 < nt && (double)(j0 + 2) / (double)nt <= sq) ++j0;          // guard the floor against rounding
    while (j0 >= 0 && (double)(j0 + 1) / (double)nt > sq) --j0;
    double m;
    const int first_knot = upper_bound(T, nt, T[0]) - 1;
    if (j0 < first_knot) {
      m = (double)T[0];                                                       // np.interp: left of the first knot -> f_0
    } else {
      const int jl = (j0 + 1 < nt && T[j0] == T[j0 + 1]) ? lower_bound(T, nt, T[j0]) - 1 : j0;   // j0 inside a run: the previous run's end
      if (jl + 1 >= nt) {
        m = (double)T[nt - 1];
      } else {
        const int jh = upper_bound(T, nt, T[jl + 1]) - 1;
        const double xl = (double)(jl + 1) / (double)nt, xh = (double)(jh + 1) / (double)nt;
        const double fl = (double)T[jl], fh = (double)T[jh];
        m = sq >= xh ? fh : fl + (sq - xl) * ((fh - fl) / (xh - xl));        // numpy: slope*(x - xp[j]) + fp[j]
      }
    }
    o[i] = (float)((1.0 - (double)weight) * (double)v + (double)weight * m);
  }
}

// ------------------------------------------------------------------------------------------------ bicubic resize, any scale factor
// F.interpolate(x, scale_factor=sf, mode='bicubic', align_corners=False) on [rows][Hi][Wi] planes (test.ipynb:553 sf = 1.25, :559 sf = 4):
// source coordinate (o + 0.5)/sf - 0.5 (ATen uses 1/scale_factor when a scale factor is given), cubic convolution A = -0.75,
// tap indices clamped to the plane.
__device__ __forceinline__ void cubic_taps(float t, float* w) {
  const float A = -0.75f;
  const float x0 = t + 1.f, x1 = t, x2 = 1.f - t, x3 = 2.f - t;
  w[0] = ((A * x0 - 5.f * A) * x0 + 8.f * A) * x0 - 4.f * A;
  w[1] = ((A + 2.f) * x1 - (A + 3.f)) * x1 * x1 + 1.f;
  w[2] = ((A + 2.f) * x2 - (A + 3.f)) * x2 * x2 + 1.f;
  w[3] = ((A * x3 - 5.f * A) * x3 + 8.f * A) * x3 - 4.f * A;
}
__global__ void __launch_bounds__(256) bicubic_resize_kernel(const float* __restrict__ x, float* __restrict__ y, long long rows, int Hi, int Wi, int Ho, int Wo,
                                                             float rh, float rw) {
  const long long total = rows * Ho * Wo;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int ox = (int)(i % Wo);
    const int oy = (int)((i / Wo) % Ho);
    const float* plane = x + (i / ((long long)Wo * Ho)) * (long long)Hi * Wi;
    const float sy = (oy + 0.5f) * rh - 0.5f, sx = (ox + 0.5f) * rw - 0.5f;
    const float fy = floorf(sy), fx = floorf(sx);
    float wy[4], wx[4];
    cubic_taps(sy - fy, wy);
    cubic_taps(sx - fx, wx);
    float acc = 0.f;
#pragma unroll
    for (int a = 0; a < 4; ++a) {
      const int yy = min(max((int)fy - 1 + a, 0), Hi - 1);
      float r = 0.f;
#pragma unroll
      for (int b = 0; b < 4; ++b) r += wx[b] * __ldg(plane + (long long)yy * Wi + min(max((int)fx - 1 + b, 0), Wi - 1));
      acc += wy[a] * r;
    }
    y[i] = acc;
  }
}

// ------------------------------------------------------------------------------------------------ feathered region blend
// smooth_blend (test.ipynb:482-496): inside the rectangle [r0, r0+h) x [c0, c0+w) out = a*(1 - m) + b*m with the feather mask m [h][w];
// outside out = a.  In place when out == a.
__global__ void __launch_bounds__(256) blend_region_kernel(const float* __restrict__ a, const float* __restrict__ b, const float* __restrict__ mask, float* __restrict__ out,
                                                           long long rows, int H, int W, int r0, int c0, int h, int w) {
  const long long total = rows * H * W;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int xx = (int)(i % W) - c0, yy = (int)((i / W) % H) - r0;
    float v = a[i];
    if (xx >= 0 && xx < w && yy >= 0 && yy < h) {
      const float m = __ldg(mask + yy * w + xx);
      v = v * (1.f - m) + __ldg(b + i) * m;
    }
    out[i] = v;
  }
}

int grid_for(long long total, int threads) {
  long long b = cdiv(total, threads);
  const long long cap = 8LL * kNumSMs;
  if (b > cap) b = cap;
  return (int)(b < 1 ? 1 : b);
}

}  // namespace

extern "C" int gdn_destandardise(const float* x, const float* trend, const unsigned char* keep, float* out, long long rows, long long HW, float scale, float shift,
                                 gdn_stream_t s) {
  GDN_CHECK_ARG(x && out && rows > 0 && HW > 0);
  destandardise_kernel<<<grid_for(rows * HW, 256), 256, 0, as_stream(s)>>>(x, trend, keep, out, rows, HW, scale, shift);
  GDN_CHECK_LAUNCH();
  return GDN_OK;
}

extern "C" int gdn_masked_spatial_mean(const float* x, const unsigned char* keep, long long rows, long long HW, float scale, float shift, float* out, gdn_stream_t s) {
  GDN_CHECK_ARG(x && out && rows > 0 && rows < (1LL << 31) && HW > 0);
  masked_spatial_mean_kernel<<<(unsigned)rows, 256, 0, as_stream(s)>>>(x, keep, HW, scale, shift, out);
  GDN_CHECK_LAUNCH();
  return GDN_OK;
}

extern "C" int gdn_ensemble_stats(const float* preds, long long member_stride, int M, long long n, float scale, float shift, float* mean, float* stdev, gdn_stream_t s) {
  GDN_CHECK_ARG(preds && mean && M > 0 && n > 0 && member_stride >= 0);
  const int grid = grid_for(n, 256);
  if (M <= 8) ensemble_stats_kernel<8><<<grid, 256, 0, as_stream(s)>>>(preds, member_stride, M, n, scale, shift, mean, stdev);
  else if (M <= 16) ensemble_stats_kernel<16><<<grid, 256, 0, as_stream(s)>>>(preds, member_stride, M, n, scale, shift, mean, stdev);
  else ensemble_stats_kernel<0><<<grid, 256, 0, as_stream(s)>>>(preds, member_stride, M, n, scale, shift, mean, stdev);
  GDN_CHECK_LAUNCH();
  return GDN_OK;
}

extern "C" int gdn_hist_match(const float* src, const float* src_sorted, const float* ref_sorted, float* out, int B, int ns, int nt, float weight, gdn_stream_t s) {
  GDN_CHECK_ARG(src && src_sorted && ref_sorted && out && B > 0 && B <= 65535 && ns > 0 && nt > 0);
  dim3 grid((unsigned)grid_for(ns, 256), (unsigned)B);
  hist_match_kernel<<<grid, 256, 0, as_stream(s)>>>(src, src_sorted, ref_sorted, out, ns, nt, weight);
  GDN_CHECK_LAUNCH();
  return GDN_OK;
}

extern "C" int gdn_bicubic_resize(const float* x, float* y, long long rows, int Hi, int Wi, int Ho, int Wo, float scale_h, float scale_w, gdn_stream_t s) {
  GDN_CHECK_ARG(x && y && rows > 0 && Hi > 0 && Wi > 0 && Ho > 0 && Wo > 0 && scale_h > 0.f && scale_w > 0.f);
  bicubic_resize_kernel<<<grid_for(rows * Ho * Wo, 256), 256, 0, as_stream(s)>>>(x, y, rows, Hi, Wi, Ho, Wo, 1.f / scale_h, 1.f / scale_w);
  GDN_CHECK_LAUNCH();
  return GDN_OK;
}

extern "C" int gdn_blend_region(const float* a, const float* b, const float* mask, float* out, long long rows, int H, int W, int r0, int c0, int h, int w, gdn_stream_t s) {
  GDN_CHECK_ARG(a && b && mask && out && rows > 0 && H > 0 && W > 0 && r0 >= 0 && c0 >= 0 && h > 0 && w > 0 && r0 + h <= H && c0 + w <= W);
  blend_region_kernel<<<grid_for(rows * H * W, 256), 256, 0, as_stream(s)>>>(a, b, mask, out, rows, H, W, r0, c0, h, w);
  GDN_CHECK_LAUNCH();
  return GDN_OK;
}
