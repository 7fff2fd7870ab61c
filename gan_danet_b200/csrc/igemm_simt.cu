// fp32 CUDA-core implicit-GEMM engine ("parity" precision).
//
//  conv_igemm_kernel : y = act(alpha * im2col(x) . W^T + bias) + res        (forward, data gradient, linear,
//                                                                             per-sample weights for CAM)
//  wgrad_igemm_kernel: out[i][j] = sum_pixels dy[m][i] * im2col(x)[m][j]     (weight gradient, Gram matrix,
//                                                                             fc1 data gradient)
// Reference call sites: nn.Conv2d / nn.Linear / torch.bmm in /root/reference/models/generator.py:20-228,
// models/discriminator.py:62-77, models/losses.py:58.  Exact fp32 FMA arithmetic, deterministic split-K.
#include "common.cuh"

namespace gdn {

struct ConvP {
  const float* x; int x_pitch, x_c0;
  const float* w; int w_k_pitch, w_c0; long long w_group_stride; int groups;
  float* y; int y_pitch, y_c0;
  const float* bias; const float* alpha_ptr;
  const float* res; int res_pitch, res_c0;
  int B, Hi, Wi, Cin, Ho, Wo, Cout, kh, kw, stride, pad, transposed;
  int act; float slope;
  int splits; float* ws;
  int Mg;       // pixels per group
  int K;        // kh*kw*Cin
  int ktiles_per_split;
};

constexpr int CBM = 128, CBK = 16, CTHREADS = 256;

__device__ __forceinline__ float conv_epilogue(float acc, float alpha, const ConvP& p, long long gm, int n) {
  float v = acc * alpha;
  if (p.bias) v += __ldg(p.bias + n);
  v = apply_act(v, p.act, p.slope);
  if (p.res) v += p.res[(size_t)gm * p.res_pitch + p.res_c0 + n];
  return v;
}

template <int BN>
__global__ void __launch_bounds__(CTHREADS) conv_igemm_kernel(const ConvP p) {
  constexpr int TN = BN / 16;
  constexpr int APITCH = CBM + 4, BPITCH = BN + 4;
  __shared__ __align__(16) float As[2][CBK][APITCH];
  __shared__ __align__(16) float Bs[2][CBK][BPITCH];

  const int t = threadIdx.x;
  const int g = blockIdx.z / p.splits, split = blockIdx.z % p.splits;
  const int m0 = blockIdx.x * CBM, n0 = blockIdx.y * BN;
  const int HoWo = p.Ho * p.Wo;
  const int taps_w = p.kw;

  // ---- loader coordinates
  const int k_local = t & 15, r_base = t >> 4;
  int a_boff[8], a_ho[8], a_wo[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    int m = m0 + r_base + 16 * i;
    if (m < p.Mg) {
      long long gm = (long long)g * p.Mg + m;
      int b = (int)(gm / HoWo), r = (int)(gm - (long long)b * HoWo);
      a_boff[i] = b * p.Hi * p.Wi;
      a_ho[i] = r / p.Wo;
      a_wo[i] = r - a_ho[i] * p.Wo;
    } else {
      a_boff[i] = 0; a_ho[i] = -(1 << 28); a_wo[i] = 0;
    }
  }
  const float* wg = p.w + (size_t)g * p.w_group_stride;
  const size_t w_row = (size_t)p.kh * p.kw * p.w_k_pitch;

  const int nk = (p.K + CBK - 1) / CBK;
  const int kt_begin = split * p.ktiles_per_split;
  const int kt_end = min(nk, kt_begin + p.ktiles_per_split);

  float a_reg[8], b_reg[TN];
  auto load_tile = [&](int kt) {
    const int k = kt * CBK + k_local;
    const bool kvalid = k < p.K;
    int tap = 0, ci = 0, fkh = 0, fkw = 0;
    if (kvalid) { tap = k / p.Cin; ci = k - tap * p.Cin; fkh = tap / taps_w; fkw = tap - fkh * taps_w; }
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      float v = 0.f;
      if (kvalid) {
        int ih, iw; bool ok;
        if (!p.transposed) {
          ih = a_ho[i] * p.stride - p.pad + fkh; iw = a_wo[i] * p.stride - p.pad + fkw;
          ok = (ih >= 0) & (ih < p.Hi) & (iw >= 0) & (iw < p.Wi);
        } else {
          int th = a_ho[i] + p.pad - fkh, tw = a_wo[i] + p.pad - fkw;
          ok = (th >= 0) & (tw >= 0);
          if (p.stride == 1) { ih = th; iw = tw; }
          else { ih = th / p.stride; iw = tw / p.stride; ok = ok & (ih * p.stride == th) & (iw * p.stride == tw); }
          ok = ok & (ih < p.Hi) & (iw < p.Wi);
        }
        if (ok) v = __ldg(p.x + (size_t)(a_boff[i] + ih * p.Wi + iw) * p.x_pitch + p.x_c0 + ci);
      }
      a_reg[i] = v;
    }
#pragma unroll
    for (int j = 0; j < TN; ++j) {
      int n = n0 + r_base + 16 * j;
      float v = 0.f;
      if (kvalid && n < p.Cout) v = __ldg(wg + (size_t)n * w_row + (size_t)tap * p.w_k_pitch + p.w_c0 + ci);
      b_reg[j] = v;
    }
  };
  auto store_tile = [&](int buf) {
#pragma unroll
    for (int i = 0; i < 8; ++i) As[buf][k_local][r_base + 16 * i] = a_reg[i];
#pragma unroll
    for (int j = 0; j < TN; ++j) Bs[buf][k_local][r_base + 16 * j] = b_reg[j];
  };

  const int tx = t & 15, ty = t >> 4;
  float acc[8][TN];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;

  if (kt_begin < kt_end) {
    load_tile(kt_begin);
    store_tile(0);
    __syncthreads();
    int buf = 0;
    for (int kt = kt_begin; kt < kt_end; ++kt) {
      const bool has_next = kt + 1 < kt_end;
      if (has_next) load_tile(kt + 1);
#pragma unroll
      for (int kk = 0; kk < CBK; ++kk) {
        float a[8], b[TN];
        *reinterpret_cast<float4*>(&a[0]) = *reinterpret_cast<const float4*>(&As[buf][kk][ty * 8]);
        *reinterpret_cast<float4*>(&a[4]) = *reinterpret_cast<const float4*>(&As[buf][kk][ty * 8 + 4]);
        if constexpr (TN == 2) {
          *reinterpret_cast<float2*>(&b[0]) = *reinterpret_cast<const float2*>(&Bs[buf][kk][tx * 2]);
        } else {
#pragma unroll
          for (int j = 0; j < TN; j += 4)
            *reinterpret_cast<float4*>(&b[j]) = *reinterpret_cast<const float4*>(&Bs[buf][kk][tx * TN + j]);
        }
#pragma unroll
        for (int i = 0; i < 8; ++i)
#pragma unroll
          for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
      }
      if (has_next) {
        store_tile(buf ^ 1);
        __syncthreads();
        buf ^= 1;
      }
    }
  }

  // ---- epilogue
  const float alpha = p.alpha_ptr ? __ldg(p.alpha_ptr) : 1.f;
  const long long Mtot = (long long)p.groups * p.Mg;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    int m = m0 + ty * 8 + i;
    if (m >= p.Mg) continue;
    long long gm = (long long)g * p.Mg + m;
#pragma unroll
    for (int j = 0; j < TN; ++j) {
      int n = n0 + tx * TN + j;
      if (n >= p.Cout) continue;
      if (p.splits > 1) {
        p.ws[((size_t)split * Mtot + gm) * p.Cout + n] = acc[i][j];
      } else {
        p.y[(size_t)gm * p.y_pitch + p.y_c0 + n] = conv_epilogue(acc[i][j], alpha, p, gm, n);
      }
    }
  }
}

__global__ void conv_splitk_epilogue_kernel(const ConvP p) {
  const long long Mtot = (long long)p.groups * p.Mg;
  const long long total = Mtot * p.Cout;
  const float alpha = p.alpha_ptr ? __ldg(p.alpha_ptr) : 1.f;
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
    long long gm = idx / p.Cout;
    int n = (int)(idx - gm * p.Cout);
    float acc = 0.f;
    for (int s = 0; s < p.splits; ++s) acc += p.ws[(size_t)s * total + idx];
    p.y[(size_t)gm * p.y_pitch + p.y_c0 + n] = conv_epilogue(acc, alpha, p, gm, n);
  }
}

// ---------------------------------------------------------------------------------------------- wgrad

struct WgradP {
  const float* dy; int dy_pitch, dy_c0;
  const float* x; int x_pitch, x_c0;
  float* out; int layout, out_cin_total, out_c0, accumulate;
  const float* scale_ptr; float scale;
  int B, Hi, Wi, Cin, Ho, Wo, Cout, kh, kw, stride, pad, groups;
  int splits; float* ws;
  int Mg, K, chunk;  // pixels per group, kh*kw*Cin, pixels per split (multiple of 16)
};

constexpr int WBI = 64, WBJ = 64, WBP = 16;

__device__ __forceinline__ void wgrad_store(const WgradP& p, int g, int i, int j, float v) {
  size_t idx;
  if (p.layout == 0) {
    idx = ((size_t)g * p.Cout + i) * p.K + j;
  } else {
    int tap = j / p.Cin, ci = j - tap * p.Cin;
    idx = ((size_t)i * p.out_cin_total + p.out_c0 + ci) * (p.kh * p.kw) + tap;
  }
  if (p.accumulate) p.out[idx] += v; else p.out[idx] = v;
}

__global__ void __launch_bounds__(256) wgrad_igemm_kernel(const WgradP p) {
  __shared__ __align__(16) float Ls[2][WBP][WBI + 4];
  __shared__ __align__(16) float Rs[2][WBP][WBJ + 4];
  const int t = threadIdx.x;
  const int g = blockIdx.z / p.splits, split = blockIdx.z % p.splits;
  const int i0 = blockIdx.x * WBI, j0 = blockIdx.y * WBJ;
  const int HoWo = p.Ho * p.Wo;

  const int c_local = t & 63, p_base = t >> 6;   // 4 pixel rows per pass, 4 passes
  const int li = i0 + c_local;
  const bool li_ok = li < p.Cout;
  const int j = j0 + c_local;
  const bool j_ok = j < p.K;
  int tap = 0, ci = 0, fkh = 0, fkw = 0;
  if (j_ok) { tap = j / p.Cin; ci = j - tap * p.Cin; fkh = tap / p.kw; fkw = tap - fkh * p.kw; }
  const bool pointwise = (p.kh == 1) & (p.kw == 1) & (p.stride == 1) & (p.pad == 0);

  const int pix_begin = split * p.chunk;
  const int pix_end = min(p.Mg, pix_begin + p.chunk);

  float l_reg[4], r_reg[4];
  auto load_tile = [&](int pb) {
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      int m = pb + p_base + 4 * r;
      float lv = 0.f, rv = 0.f;
      if (m < pix_end) {
        long long gm = (long long)g * p.Mg + m;
        if (li_ok) lv = __ldg(p.dy + (size_t)gm * p.dy_pitch + p.dy_c0 + li);
        if (j_ok) {
          if (pointwise) {
            rv = __ldg(p.x + (size_t)gm * p.x_pitch + p.x_c0 + ci);
          } else {
            int b = (int)(gm / HoWo), rr = (int)(gm - (long long)b * HoWo);
            int ho = rr / p.Wo, wo = rr - ho * p.Wo;
            int ih = ho * p.stride - p.pad + fkh, iw = wo * p.stride - p.pad + fkw;
            if ((ih >= 0) & (ih < p.Hi) & (iw >= 0) & (iw < p.Wi))
              rv = __ldg(p.x + ((size_t)b * p.Hi * p.Wi + (size_t)ih * p.Wi + iw) * p.x_pitch + p.x_c0 + ci);
          }
        }
      }
      l_reg[r] = lv; r_reg[r] = rv;
    }
  };
  auto store_tile = [&](int buf) {
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      Ls[buf][p_base + 4 * r][c_local] = l_reg[r];
      Rs[buf][p_base + 4 * r][c_local] = r_reg[r];
    }
  };

  const int tx = t & 15, ty = t >> 4;
  float acc[4][4];
#pragma unroll
  for (int a = 0; a < 4; ++a)
#pragma unroll
    for (int b = 0; b < 4; ++b) acc[a][b] = 0.f;

  if (pix_begin < pix_end) {
    load_tile(pix_begin);
    store_tile(0);
    __syncthreads();
    int buf = 0;
    for (int pb = pix_begin; pb < pix_end; pb += WBP) {
      const bool has_next = pb + WBP < pix_end;
      if (has_next) load_tile(pb + WBP);
#pragma unroll
      for (int kk = 0; kk < WBP; ++kk) {
        float4 a4 = *reinterpret_cast<const float4*>(&Ls[buf][kk][ty * 4]);
        float4 b4 = *reinterpret_cast<const float4*>(&Rs[buf][kk][tx * 4]);
        float a[4] = {a4.x, a4.y, a4.z, a4.w}, b[4] = {b4.x, b4.y, b4.z, b4.w};
#pragma unroll
        for (int u = 0; u < 4; ++u)
#pragma unroll
          for (int v = 0; v < 4; ++v) acc[u][v] = fmaf(a[u], b[v], acc[u][v]);
      }
      if (has_next) {
        store_tile(buf ^ 1);
        __syncthreads();
        buf ^= 1;
      }
    }
  }

  const float scale = p.scale * (p.scale_ptr ? __ldg(p.scale_ptr) : 1.f);
#pragma unroll
  for (int u = 0; u < 4; ++u) {
    int i = i0 + ty * 4 + u;
    if (i >= p.Cout) continue;
#pragma unroll
    for (int v = 0; v < 4; ++v) {
      int jj = j0 + tx * 4 + v;
      if (jj >= p.K) continue;
      if (p.splits > 1) p.ws[(((size_t)g * p.splits + split) * p.Cout + i) * p.K + jj] = acc[u][v];
      else wgrad_store(p, g, i, jj, acc[u][v] * scale);
    }
  }
}

__global__ void wgrad_reduce_kernel(const WgradP p) {
  const long long per_group = (long long)p.Cout * p.K;
  const long long total = per_group * p.groups;
  const float scale = p.scale * (p.scale_ptr ? __ldg(p.scale_ptr) : 1.f);
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
    int g = (int)(idx / per_group);
    long long r = idx - (long long)g * per_group;
    int i = (int)(r / p.K), j = (int)(r - (long long)i * p.K);
    float acc = 0.f;
    for (int s = 0; s < p.splits; ++s) acc += p.ws[((size_t)g * p.splits + s) * per_group + r];
    wgrad_store(p, g, i, j, acc * scale);
  }
}

}  // namespace gdn

using namespace gdn;

static int conv_tiles(const gdn_conv_args* a, int* bn_out) {
  int groups = a->groups > 0 ? a->groups : 1;
  long long Mg = (long long)(a->B / groups) * a->Ho * a->Wo;
  int bn = a->Cout <= 32 ? 32 : (a->Cout <= 64 ? 64 : 128);
  if (bn == 128 && cdiv(Mg, CBM) * cdiv(a->Cout, 128) * groups < 2 * kNumSMs && a->Cout % 128 != 0 && a->Cout % 64 == 0) bn = 64;
  if (bn_out) *bn_out = bn;
  return (int)(cdiv(Mg, CBM) * cdiv(a->Cout, bn) * groups);
}

extern "C" int gdn_conv2d_suggest_splits(const gdn_conv_args* a) {
  int tiles = conv_tiles(a, nullptr);
  int nk = (int)cdiv((long long)a->kh * a->kw * a->Cin, CBK);
  if (tiles >= kNumSMs) return 1;
  int s = (int)cdiv(2 * kNumSMs, tiles);
  int max_s = nk / 8 > 0 ? nk / 8 : 1;
  if (s > max_s) s = max_s;
  if (s > 64) s = 64;
  return s < 1 ? 1 : s;
}

extern "C" int gdn_conv2d(const gdn_conv_args* a, gdn_stream_t s) {
  GDN_CHECK_ARG(a && a->x && a->w && a->y);
  GDN_CHECK_ARG(a->B > 0 && a->Cin > 0 && a->Cout > 0 && a->kh > 0 && a->kw > 0 && a->stride > 0);
  GDN_CHECK_ARG(a->Hi > 0 && a->Wi > 0 && a->Ho > 0 && a->Wo > 0);
  int groups = a->groups > 0 ? a->groups : 1;
  GDN_CHECK_ARG(a->B % groups == 0);
  GDN_CHECK_ARG(a->x_pitch >= a->x_c0 + a->Cin && a->y_pitch >= a->y_c0 + a->Cout && a->w_k_pitch >= a->w_c0 + a->Cin);
  GDN_CHECK_ARG((long long)a->B * a->Hi * a->Wi < (1ll << 31) && (long long)a->B * a->Ho * a->Wo < (1ll << 31));
  ConvP p;
  p.x = a->x; p.x_pitch = a->x_pitch; p.x_c0 = a->x_c0;
  p.w = a->w; p.w_k_pitch = a->w_k_pitch; p.w_c0 = a->w_c0; p.w_group_stride = a->w_group_stride; p.groups = groups;
  p.y = a->y; p.y_pitch = a->y_pitch; p.y_c0 = a->y_c0;
  p.bias = a->bias; p.alpha_ptr = a->alpha_ptr; p.res = a->res; p.res_pitch = a->res_pitch; p.res_c0 = a->res_c0;
  p.B = a->B; p.Hi = a->Hi; p.Wi = a->Wi; p.Cin = a->Cin; p.Ho = a->Ho; p.Wo = a->Wo; p.Cout = a->Cout;
  p.kh = a->kh; p.kw = a->kw; p.stride = a->stride; p.pad = a->pad; p.transposed = a->transposed;
  p.act = a->act; p.slope = a->slope;
  p.splits = a->splits > 1 ? a->splits : 1; p.ws = a->ws;
  p.Mg = (a->B / groups) * a->Ho * a->Wo;
  p.K = a->kh * a->kw * a->Cin;
  int nk = (int)cdiv(p.K, CBK);
  if (p.splits > nk) p.splits = nk;
  p.ktiles_per_split = (int)cdiv(nk, p.splits);
  p.splits = (int)cdiv(nk, p.ktiles_per_split);
  long long Mtot = (long long)groups * p.Mg;
  if (p.splits > 1) {
    GDN_CHECK_ARG(a->ws != nullptr);
    if (a->ws_bytes < (size_t)p.splits * Mtot * p.Cout * sizeof(float)) { set_error("gdn_conv2d: workspace too small"); return GDN_EWORKSPACE; }
  }
  int bn;
  conv_tiles(a, &bn);
  dim3 grid((unsigned)cdiv(p.Mg, CBM), (unsigned)cdiv(p.Cout, bn), (unsigned)(groups * p.splits));
  GDN_CHECK_ARG(grid.y <= 65535 && grid.z <= 65535);
  cudaStream_t st = as_stream(s);
  if (bn == 32) conv_igemm_kernel<32><<<grid, CTHREADS, 0, st>>>(p);
  else if (bn == 64) conv_igemm_kernel<64><<<grid, CTHREADS, 0, st>>>(p);
  else conv_igemm_kernel<128><<<grid, CTHREADS, 0, st>>>(p);
  GDN_CHECK_LAUNCH();
  if (p.splits > 1) {
    long long total = Mtot * p.Cout;
    int blocks = (int)(cdiv(total, 256) < 4 * kNumSMs ? cdiv(total, 256) : 4 * kNumSMs);
    conv_splitk_epilogue_kernel<<<blocks, 256, 0, st>>>(p);
    GDN_CHECK_LAUNCH();
  }
  return GDN_OK;
}

extern "C" int gdn_wgrad_suggest_splits(const gdn_wgrad_args* a) {
  int groups = a->groups > 0 ? a->groups : 1;
  long long K = (long long)a->kh * a->kw * a->Cin;
  long long tiles = cdiv(a->Cout, WBI) * cdiv(K, WBJ) * groups;
  long long Mg = (long long)(a->B / groups) * a->Ho * a->Wo;
  long long s = cdiv(3 * kNumSMs, tiles);
  long long max_s = Mg / 256 > 0 ? Mg / 256 : 1;
  if (s > max_s) s = max_s;
  if (s > 512) s = 512;
  return (int)(s < 1 ? 1 : s);
}

extern "C" int gdn_conv2d_wgrad(const gdn_wgrad_args* a, gdn_stream_t s) {
  GDN_CHECK_ARG(a && a->dy && a->x && a->out);
  GDN_CHECK_ARG(a->B > 0 && a->Cin > 0 && a->Cout > 0 && a->kh > 0 && a->kw > 0 && a->stride > 0);
  int groups = a->groups > 0 ? a->groups : 1;
  GDN_CHECK_ARG(a->B % groups == 0);
  GDN_CHECK_ARG(a->layout == 0 || (a->layout == 1 && groups == 1 && a->out_cin_total >= a->out_c0 + a->Cin));
  GDN_CHECK_ARG(a->x_pitch >= a->x_c0 + a->Cin && a->dy_pitch >= a->dy_c0 + a->Cout);
  WgradP p;
  p.dy = a->dy; p.dy_pitch = a->dy_pitch; p.dy_c0 = a->dy_c0;
  p.x = a->x; p.x_pitch = a->x_pitch; p.x_c0 = a->x_c0;
  p.out = a->out; p.layout = a->layout; p.out_cin_total = a->out_cin_total; p.out_c0 = a->out_c0; p.accumulate = a->accumulate;
  p.scale_ptr = a->scale_ptr; p.scale = a->scale;
  p.B = a->B; p.Hi = a->Hi; p.Wi = a->Wi; p.Cin = a->Cin; p.Ho = a->Ho; p.Wo = a->Wo; p.Cout = a->Cout;
  p.kh = a->kh; p.kw = a->kw; p.stride = a->stride; p.pad = a->pad; p.groups = groups;
  p.Mg = (a->B / groups) * a->Ho * a->Wo;
  p.K = a->kh * a->kw * a->Cin;
  p.splits = a->splits > 1 ? a->splits : 1;
  p.chunk = (int)(cdiv(cdiv(p.Mg, p.splits), WBP) * WBP);
  p.splits = (int)cdiv(p.Mg, p.chunk);
  p.ws = a->ws;
  if (p.splits > 1) {
    GDN_CHECK_ARG(a->ws != nullptr);
    if (a->ws_bytes < (size_t)p.splits * groups * p.Cout * (size_t)p.K * sizeof(float)) { set_error("gdn_conv2d_wgrad: workspace too small"); return GDN_EWORKSPACE; }
  }
  dim3 grid((unsigned)cdiv(p.Cout, WBI), (unsigned)cdiv(p.K, WBJ), (unsigned)(groups * p.splits));
  GDN_CHECK_ARG(grid.y <= 65535 && grid.z <= 65535);
  cudaStream_t st = as_stream(s);
  wgrad_igemm_kernel<<<grid, 256, 0, st>>>(p);
  GDN_CHECK_LAUNCH();
  if (p.splits > 1) {
    long long total = (long long)groups * p.Cout * p.K;
    int blocks = (int)(cdiv(total, 256) < 4 * kNumSMs ? cdiv(total, 256) : 4 * kNumSMs);
    wgrad_reduce_kernel<<<blocks, 256, 0, st>>>(p);
    GDN_CHECK_LAUNCH();
  }
  return GDN_OK;
}
