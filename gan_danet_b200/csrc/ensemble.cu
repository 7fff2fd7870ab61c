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
template <int V>
__global__ void __launch_bounds__(256) destandardise_kernel(const float* __restrict__ x, const float* __restrict__ trend, const unsigned char* __restrict__ keep,
                                                            float* __restrict__ out, long long rows, long long HW, float scale, float shift) {
  const long long total = rows * HW / V;
  const float nanv = __int_as_float(0x7fc00000);
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    float v[V];
    if (V == 4) {
      const float4 a = __ldg(reinterpret_cast<const float4*>(x) + i);
      v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w;
      if (trend) {
        const float4 t = __ldg(reinterpret_cast<const float4*>(trend) + i);
        v[0] += t.x; v[1] += t.y; v[2] += t.z; v[3] += t.w;
      }
    } else {
      v[0] = __ldg(x + i);
      if (trend) v[0] += __ldg(trend + i);
    }
    const long long p0 = (i * V) % HW;                  // V divides HW on the vector path: the V pixels are in one field
#pragma unroll
    for (int e = 0; e < V; ++e) {
      v[e] = fmaf(v[e], scale, shift);
      if (keep && !keep[p0 + e]) v[e] = nanv;
    }
    if (V == 4) reinterpret_cast<float4*>(out)[i] = make_float4(v[0], v[1], v[2], v[3]);
    else out[i] = v[0];
  }
}

// ------------------------------------------------------------------------------------------------ masked spatial mean
// out[r] = nanmean_p { x[r][p]*scale + shift : keep[p] != 0 }  (NaN inputs are skipped like np.nanmean; no valid pixel -> NaN).
// Grid (rows, splits): every CTA reduces one contiguous chunk of one row (128-bit loads when the chunk allows, per-thread partial sums in
// double, fixed-order block reduction) into ws[row][split] = (sum, count); the finish kernel adds the splits in order.  Bitwise reproducible.
__global__ void __launch_bounds__(256) masked_spatial_mean_kernel(const float* __restrict__ x, const unsigned char* __restrict__ keep, long long HW, long long chunk,
                                                                  float scale, float shift, double2* __restrict__ ws, int vec4) {
  __shared__ double red[32];
  const float* row = x + (long long)blockIdx.x * HW;
  const long long p0 = (long long)blockIdx.y * chunk;
  const long long p1 = p0 + chunk < HW ? p0 + chunk : HW;
  double s = 0.0, c = 0.0;
  if (vec4) {
    for (long long p = p0 + 4LL * threadIdx.x; p < p1; p += 4LL * blockDim.x) {
      const float4 v = __ldg(reinterpret_cast<const float4*>(row + p));
      const float f[4] = {v.x, v.y, v.z, v.w};
      unsigned k4 = 0x01010101u;
      if (keep) k4 = __ldg(reinterpret_cast<const unsigned*>(keep + p));
#pragma unroll
      for (int e = 0; e < 4; ++e)
        if (((k4 >> (8 * e)) & 0xffu) && f[e] == f[e]) { s += (double)fmaf(f[e], scale, shift); c += 1.0; }
    }
  } else {
    for (long long p = p0 + threadIdx.x; p < p1; p += blockDim.x) {
      const float v = __ldg(row + p);
      if ((keep == nullptr || keep[p]) && v == v) { s += (double)fmaf(v, scale, shift); c += 1.0; }
    }
  }
  s = block_sum(s, red);
  c = block_sum(c, red);
  if (threadIdx.x == 0) ws[(long long)blockIdx.x * gridDim.y + blockIdx.y] = make_double2(s, c);
}
__global__ void masked_spatial_mean_finish_kernel(const double2* __restrict__ ws, int splits, long long rows, float* __restrict__ out) {
  const long long r = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (r >= rows) return;
  double s = 0.0, c = 0.0;
  for (int k = 0; k < splits; ++k) { const double2 v = ws[r * splits + k]; s += v.x; c += v.y; }
  out[r] = c > 0.0 ? (float)(s / c) : __int_as_float(0x7fc00000);
}

// ------------------------------------------------------------------------------------------------ statistics over the members
// preds: M fields of n elements (member m at preds + m*stride).  mean[i], std[i] over the non-NaN members (ddof = 0, np.nanstd);
// two passes over the members held in registers for M <= 16 (the one-per-GPU ensemble), re-read otherwise.
template <int MAXM, int V>
__global__ void __launch_bounds__(256) ensemble_stats_kernel(const float* __restrict__ preds, long long stride, int M, long long n, float scale, float shift,
                                                             float* __restrict__ mean, float* __restrict__ stdev) {
  const float nanv = __int_as_float(0x7fc00000);
  const long long total = n / V;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    float v[MAXM > 0 ? MAXM : 1][V];
    double s[V];
    int c[V];
#pragma unroll
    for (int e = 0; e < V; ++e) { s[e] = 0.0; c[e] = 0; }
    auto load = [&](int m, float* dst) {
      if (V == 4) {
        const float4 f = __ldg(reinterpret_cast<const float4*>(preds + m * stride) + i);
        dst[0] = f.x; dst[1] = f.y; dst[2] = f.z; dst[3] = f.w;
      } else {
        dst[0] = __ldg(preds + m * stride + i);
      }
#pragma unroll
      for (int e = 0; e < V; ++e) dst[e] = fmaf(dst[e], scale, shift);
    };
    if (MAXM > 0) {
#pragma unroll
      for (int m = 0; m < MAXM; ++m) {
        if (m < M) load(m, v[m]);
        else {
#pragma unroll
          for (int e = 0; e < V; ++e) v[m][e] = nanv;
        }
#pragma unroll
        for (int e = 0; e < V; ++e)
          if (v[m][e] == v[m][e]) { s[e] += (double)v[m][e]; ++c[e]; }
      }
    } else {
      for (int m = 0; m < M; ++m) {
        float t[V];
        load(m, t);
#pragma unroll
        for (int e = 0; e < V; ++e)
          if (t[e] == t[e]) { s[e] += (double)t[e]; ++c[e]; }
      }
    }
    double mu[V], q[V];
#pragma unroll
    for (int e = 0; e < V; ++e) { mu[e] = c[e] ? s[e] / c[e] : 0.0; q[e] = 0.0; }
    if (stdev) {
      if (MAXM > 0) {
#pragma unroll
        for (int m = 0; m < MAXM; ++m)
#pragma unroll
          for (int e = 0; e < V; ++e)
            if (v[m][e] == v[m][e]) { const double dlt = (double)v[m][e] - mu[e]; q[e] += dlt * dlt; }
      } else {
        for (int m = 0; m < M; ++m) {
          float t[V];
          load(m, t);
#pragma unroll
          for (int e = 0; e < V; ++e)
            if (t[e] == t[e]) { const double dlt = (double)t[e] - mu[e]; q[e] += dlt * dlt; }
        }
      }
    }
    float om[V], os[V];
#pragma unroll
    for (int e = 0; e < V; ++e) { om[e] = c[e] ? (float)mu[e] : nanv; os[e] = c[e] ? (float)sqrt(q[e] / c[e]) : nanv; }
    if (V == 4) {
      reinterpret_cast<float4*>(mean)[i] = make_float4(om[0], om[1], om[2], om[3]);
      if (stdev) reinterpret_cast<float4*>(stdev)[i] = make_float4(os[0], os[1], os[2], os[3]);
    } else {
      mean[i] = om[0];
      if (stdev) stdev[i] = os[0];
    }
  }
}

// ------------------------------------------------------------------------------------------------ row-wise sort (np.sort / np.unique of test.ipynb:117-121)
// Ascending sort of every row of [rows][n] (NaNs last, like np.sort), one CTA per row: bitonic network on order-preserving 32-bit keys.  Stages whose
// compare distance fits a shared-memory chunk (8192 keys) run in shared memory, the log2(np / 8192) largest distances of each merge run on the row's
// L2-resident key buffer in `ws`.  Fixed network => bitwise reproducible; replaces the two library sorts that were 4.4 of the 6 ms of histogram matching.
constexpr int SORT_CH = 8192;
__device__ __forceinline__ unsigned sort_key(float f) {
  if (f != f) return 0xFFFFFFFEu;                                   // every NaN after every number (padding is 0xFFFFFFFF)
  const unsigned u = __float_as_uint(f);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float sort_unkey(unsigned k) {
  if (k >= 0xFFFFFFFEu) return __int_as_float(0x7fc00000);
  return __uint_as_float((k & 0x80000000u) ? (k & 0x7FFFFFFFu) : ~k);
}
__device__ __forceinline__ void cmp_swap(unsigned& a, unsigned& b, bool ascending) {
  if ((a > b) == ascending) { const unsigned t = a; a = b; b = t; }
}
__global__ void __launch_bounds__(1024, 2) sort_rows_kernel(const float* __restrict__ x, float* __restrict__ out, unsigned* __restrict__ ws, int n, int np) {
  __shared__ unsigned sk[SORT_CH];
  const int tid = threadIdx.x;
  unsigned* g = ws + (size_t)blockIdx.x * np;
  const float* xr = x + (size_t)blockIdx.x * n;
  for (int i = tid; i < np; i += 1024) g[i] = i < n ? sort_key(xr[i]) : 0xFFFFFFFFu;
  __syncthreads();
  const int ch = np < SORT_CH ? np : SORT_CH;
  // all merge sizes k <= ch, chunk by chunk in shared memory; then for every larger k: the distances >= ch in global memory, the rest per chunk
  for (int k0 = 2; k0 <= np; k0 = (k0 < ch ? ch : k0) << 1) {
    const int kfirst = k0 <= ch ? 2 : k0, klast = k0 <= ch ? ch : k0;      // first round: k = 2 .. ch inside the chunks
    if (k0 > ch) {
      for (int j = k0 >> 1; j >= ch; j >>= 1) {
        for (int t = tid; t < (np >> 1); t += 1024) {
          const int i = ((t & ~(j - 1)) << 1) | (t & (j - 1));
          unsigned a = g[i], b = g[i | j];
          cmp_swap(a, b, (i & k0) == 0);
          g[i] = a; g[i | j] = b;
        }
        __syncthreads();
      }
    }
    for (int c0 = 0; c0 < np; c0 += ch) {
      for (int i = tid; i < ch; i += 1024) sk[i] = g[c0 + i];
      __syncthreads();
      for (int k = kfirst; k <= klast; k <<= 1) {
        for (int j = (k > ch ? ch : k) >> 1; j >= 1; j >>= 1) {
          for (int t = tid; t < (ch >> 1); t += 1024) {
            const int i = ((t & ~(j - 1)) << 1) | (t & (j - 1));
            unsigned a = sk[i], b = sk[i | j];
            cmp_swap(a, b, ((c0 + i) & k) == 0);
            sk[i] = a; sk[i | j] = b;
          }
          __syncthreads();
        }
      }
      for (int i = tid; i < ch; i += 1024) g[c0 + i] = sk[i];
      __syncthreads();
    }
  }
  float* o = out + (size_t)blockIdx.x * n;
  for (int i = tid; i < n; i += 1024) o[i] = sort_unkey(g[i]);
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
  __shared__ int first_knot_s;                                                  // end of the first run of T: once per CTA, not per element
  if (threadIdx.x == 0) first_knot_s = upper_bound(T, nt, T[0]) - 1;
  __syncthreads();
  const int first_knot = first_knot_s;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < ns; i += gridDim.x * blockDim.x) {
    const float v = x[i];
    const double sq = (double)upper_bound(S, ns, v) / (double)ns;
    // largest knot j with (j+1)/nt <= sq
    int j0 = (int)floor(sq * (double)nt) - 1;
    while (j0 + 1 < nt && (double)(j0 + 2) / (double)nt <= sq) ++j0;          // guard the floor against rounding
    while (j0 >= 0 && (double)(j0 + 1) / (double)nt > sq) --j0;
    double m;
    if (j0 < first_knot) {
      m = (double)T[0];                                                       // np.interp: left of the first knot -> f_0
    } else {
      // j0 inside a run of equal values: the previous run's end.  Distinct neighbours (the common case for continuous fields) need no search.
      const int jl = (j0 + 1 < nt && T[j0] == T[j0 + 1]) ? lower_bound(T, nt, T[j0]) - 1 : j0;
      if (jl + 1 >= nt) {
        m = (double)T[nt - 1];
      } else {
        const int jh = (jl + 2 < nt && T[jl + 1] == T[jl + 2]) ? upper_bound(T, nt, T[jl + 1]) - 1 : jl + 1;
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
  const long long cap = 32LL * kNumSMs;        // 8 resident CTAs of 256 threads per SM x 4 waves of grid-stride work
  if (b > cap) b = cap;
  return (int)(b < 1 ? 1 : b);
}

}  // namespace

extern "C" int gdn_destandardise(const float* x, const float* trend, const unsigned char* keep, float* out, long long rows, long long HW, float scale, float shift,
                                 gdn_stream_t s) {
  GDN_CHECK_ARG(x && out && rows > 0 && HW > 0);
  const bool v4 = HW % 4 == 0 && (((uintptr_t)x | (uintptr_t)out | (uintptr_t)trend) & 15) == 0;
  if (v4) destandardise_kernel<4><<<grid_for(rows * HW / 4, 256), 256, 0, as_stream(s)>>>(x, trend, keep, out, rows, HW, scale, shift);
  else destandardise_kernel<1><<<grid_for(rows * HW, 256), 256, 0, as_stream(s)>>>(x, trend, keep, out, rows, HW, scale, shift);
  GDN_CHECK_LAUNCH();
  return GDN_OK;
}

static int masked_mean_splits(long long rows, long long HW) {
  long long want = cdiv(8LL * kNumSMs, rows);                  // >= 8 CTAs per SM in flight
  const long long most = cdiv(HW, 4096);                       // a chunk is at least 4096 pixels (16 per thread)
  if (want > most) want = most;
  return (int)(want < 1 ? 1 : want);
}
extern "C" size_t gdn_masked_spatial_mean_ws_bytes(long long rows, long long HW) { return (size_t)rows * masked_mean_splits(rows, HW) * sizeof(double2); }
extern "C" int gdn_masked_spatial_mean(const float* x, const unsigned char* keep, long long rows, long long HW, float scale, float shift, float* out, void* ws,
                                       size_t ws_bytes, gdn_stream_t s) {
  GDN_CHECK_ARG(x && out && ws && rows > 0 && rows < (1LL << 31) && HW > 0 && ((uintptr_t)ws & 15) == 0);
  if (ws_bytes < gdn_masked_spatial_mean_ws_bytes(rows, HW)) { set_error("gdn_masked_spatial_mean: workspace too small"); return GDN_EWORKSPACE; }
  const int splits = masked_mean_splits(rows, HW);
  long long chunk = cdiv(HW, splits);
  chunk = (chunk + 3) & ~3LL;
  const int vec4 = HW % 4 == 0 && ((uintptr_t)x & 15) == 0 && (keep == nullptr || ((uintptr_t)keep & 3) == 0);
  dim3 grid((unsigned)rows, (unsigned)splits);
  GDN_CHECK_ARG(splits <= 65535);
  masked_spatial_mean_kernel<<<grid, 256, 0, as_stream(s)>>>(x, keep, HW, chunk, scale, shift, reinterpret_cast<double2*>(ws), vec4);
  GDN_CHECK_LAUNCH();
  masked_spatial_mean_finish_kernel<<<(unsigned)cdiv(rows, 256), 256, 0, as_stream(s)>>>(reinterpret_cast<const double2*>(ws), splits, rows, out);
  GDN_CHECK_LAUNCH();
  return GDN_OK;
}

extern "C" int gdn_ensemble_stats(const float* preds, long long member_stride, int M, long long n, float scale, float shift, float* mean, float* stdev, gdn_stream_t s) {
  GDN_CHECK_ARG(preds && mean && M > 0 && n > 0 && member_stride >= 0);
  const bool v4 = n % 4 == 0 && member_stride % 4 == 0 && (((uintptr_t)preds | (uintptr_t)mean | (uintptr_t)stdev) & 15) == 0;
  cudaStream_t st = as_stream(s);
  if (v4) {
    const int grid = grid_for(n / 4, 256);
    if (M <= 8) ensemble_stats_kernel<8, 4><<<grid, 256, 0, st>>>(preds, member_stride, M, n, scale, shift, mean, stdev);
    else ensemble_stats_kernel<0, 4><<<grid, 256, 0, st>>>(preds, member_stride, M, n, scale, shift, mean, stdev);
  } else {
    const int grid = grid_for(n, 256);
    if (M <= 8) ensemble_stats_kernel<8, 1><<<grid, 256, 0, st>>>(preds, member_stride, M, n, scale, shift, mean, stdev);
    else if (M <= 16) ensemble_stats_kernel<16, 1><<<grid, 256, 0, st>>>(preds, member_stride, M, n, scale, shift, mean, stdev);
    else ensemble_stats_kernel<0, 1><<<grid, 256, 0, st>>>(preds, member_stride, M, n, scale, shift, mean, stdev);
  }
  GDN_CHECK_LAUNCH();
  return GDN_OK;
}

static int next_pow2(int v) { int p = 1; while (p < v) p <<= 1; return p; }
extern "C" size_t gdn_sort_rows_ws_bytes(int rows, int n) { return (size_t)rows * (size_t)next_pow2(n > 1 ? n : 2) * sizeof(unsigned); }
extern "C" int gdn_sort_rows(const float* x, float* out, int rows, int n, void* ws, size_t ws_bytes, gdn_stream_t s) {
  GDN_CHECK_ARG(x && out && rows > 0 && n > 0 && n <= (1 << 26));
  if (!ws || ws_bytes < gdn_sort_rows_ws_bytes(rows, n)) { set_error("gdn_sort_rows: workspace too small"); return GDN_EWORKSPACE; }
  sort_rows_kernel<<<rows, 1024, 0, as_stream(s)>>>(x, out, reinterpret_cast<unsigned*>(ws), n, next_pow2(n > 1 ? n : 2));
  GDN_CHECK_LAUNCH();
  return GDN_OK;
}

extern "C" int gdn_hist_match(const float* src, const float* src_sorted, const float* ref_sorted, float* out, int B, int ns, int nt, float weight, gdn_stream_t s) {
  GDN_CHECK_ARG(src && src_sorted && ref_sorted && out && B > 0 && B <= 65535 && ns > 0 && nt > 0);
  long long per_sample = cdiv(16LL * kNumSMs, B);                  // ~16 CTAs per SM over all samples, at least 4 elements per thread
  const long long most = cdiv(ns, 1024);
  if (per_sample > most) per_sample = most;
  dim3 grid((unsigned)(per_sample < 1 ? 1 : per_sample), (unsigned)B);
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
