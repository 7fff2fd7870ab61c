// nn.Linear on the tensor cores in TF32, for Discriminator1.fc1 (/root/reference/models/discriminator.py:66,75): a
// [B, 262144] x [262144 -> 1024] layer whose 1 GB fp32 weight makes all three GEMMs HBM-bound.  The fp32
// weight is consumed AS IT IS by tcgen05.mma kind::tf32 through TMA (no packed copy exists: any conversion pass would double the
// traffic of a layer that is pure weight streaming); the skinny operand (activations / output gradient, <= 64 rows) rides along.
//   forward   y[m][n]  = act(sum_k x[m][k] W[n][k] + b[n]) : D[n][m], A = W tile (K-major), B = x tile (K-major), split-K over CTAs
//   dgrad     dx[m][k] = sum_n dz[m][n] W[n][k]            : D[k][m], A = W tile read MN-major (k contiguous), B = dz (K-major)
//   wgrad     dW[n][k] = sum_m dz[m][n] x[m][k]            : D[k][n], A = x read MN-major, B = dz^T [n][m] (K-major; a 256 KB transpose), K = batch rows
//             (both gradients read a 32-bit operand MN-major: this needs the 128B-swizzle-with-32-byte-atom layout, see desc_mn32)
// All three: persistent CTAs, warp 0 TMA producer, warp 1 MMA issuer, warp 2 TMEM allocator, warps 4-7 epilogue, two accumulator
// buffers in TMEM (the epilogue of one tile overlaps the main loop of the next).
#include "tc_common.cuh"

namespace gdn {
namespace lintc {
using namespace gdn::tc;

constexpr int NT = 256;
constexpr int SMEM_LIMIT = 227 * 1024;
constexpr int BOXK = 32;                       // fp32 elements per 128-byte swizzled row
enum { MODE_FWD = 0, MODE_DGRAD = 1, MODE_WGRAD = 2 };

// kind::tf32 instruction descriptor: D fp32 (bit 4), A/B tf32 (2 at [7,10), [10,13)), a_major bit 15, b_major bit 16 (1 = MN-major)
__host__ __device__ constexpr uint32_t idesc(int M, int N, int a_mn, int b_mn) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)a_mn << 15) | ((uint32_t)b_mn << 16) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
// MN-major operand of a 32-bit type: the transposing read needs the "128B swizzle with 32-byte atomicity" layout (UMMA layout type 1,
// cute's Layout_MN_SW128_32B_Atom = Swizzle<2,5,2>, 4 K-rows per atom; TMA mode CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B).  Measured on B200:
// with the plain 128B swizzle (layout type 2) kind::tf32 returns zeros for an MN-major operand.
constexpr uint32_t LAYOUT_SW128_BASE32B = 1;
__device__ __forceinline__ uint64_t desc_mn32(uint32_t addr, uint32_t lbo_bytes) {     // one MMA step = 8 K-rows = two 4-row atoms 512 B apart
  return (uint64_t)((addr >> 4) & 0x3FFF) | ((uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16) | ((uint64_t)((512u >> 4) & 0x3FFF) << 32) | ((uint64_t)1 << 46) |
         ((uint64_t)LAYOUT_SW128_BASE32B << 61);
}

struct Params {
  int Mb;            // batch rows (multiple of 16, <= 64)
  int N, K;          // out features, in features
  int stages, stage_bytes, a_bytes;
  int tiles, ksteps; // persistent tile count; pipeline stages per tile
  int ksplit;        // forward: K splits
  float* out;        // forward: partial sums [ksplit][N][Mb]; dgrad: dx [Mb][K]; wgrad: dW [N][K]
};

template <int MODE>
__global__ void __launch_bounds__(NT, 1)
linear_tc_kernel(const __grid_constant__ CUtensorMap mapW, const __grid_constant__ CUtensorMap mapX, const __grid_constant__ CUtensorMap mapDZ, const Params p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* sm = smem_raw + (base - smem_u32(smem_raw));
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  constexpr int MAXS = 8;
  const uint32_t epi_off = p.stages * p.stage_bytes;                 // 4 epilogue warps x [32][32] fp32 transpose stages
  const uint32_t bar_base = base + epi_off + 4 * 4096;
  auto full = [&](int s) { return bar_base + 8 * s; };
  auto empty = [&](int s) { return bar_base + 8 * (MAXS + s); };
  auto acc_full = [&](int b) { return bar_base + 8 * (2 * MAXS + b); };
  auto acc_empty = [&](int b) { return bar_base + 8 * (2 * MAXS + 2 + b); };
  const uint32_t tslot = bar_base + 8 * (2 * MAXS + 4);
  const int NW = MODE == MODE_WGRAD ? 256 : p.Mb;                    // accumulator width (UMMA N)
  const uint32_t tmem_cols = MODE == MODE_WGRAD ? 512u : 128u;
  const uint32_t acc_stride = tmem_cols / 2;

  if (threadIdx.x == 0) {
    for (int s = 0; s < p.stages; ++s) { mbar_init(full(s), 1); mbar_init(empty(s), 1); }
    for (int b = 0; b < 2; ++b) { mbar_init(acc_full(b), 1); mbar_init(acc_empty(b), 4); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tslot), "r"(tmem_cols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *reinterpret_cast<volatile uint32_t*>(sm + (tslot - base));
  const int row_bytes = p.Mb * 128;          // one [Mb rows][32 fp32] box

  if (warp == 0) {
    if (lane == 0) {   // ---- TMA producer
      int it = 0;
      for (int tile = blockIdx.x; tile < p.tiles; tile += gridDim.x) {
        for (int ks = 0; ks < p.ksteps; ++ks, ++it) {
          const int s = it % p.stages;
          if (it >= p.stages) mbar_wait(empty(s), ((it / p.stages) - 1) & 1);
          const uint32_t a = base + s * p.stage_bytes, b = a + p.a_bytes;
          mbar_expect_tx(full(s), p.stage_bytes);
          if (MODE == MODE_FWD) {            // tile = (n block, K split); stage = 64 k: 2 W boxes [128 n][32 k] + 2 x boxes [Mb][32 k]
            const int nb = tile / p.ksplit, sp = tile % p.ksplit;
            const int k0 = (sp * p.ksteps + ks) * 64;
            for (int j = 0; j < 2; ++j) {
              tma_load_2d(a + j * 16384, &mapW, full(s), k0 + j * BOXK, nb * 128);
              tma_load_2d(b + j * row_bytes, &mapX, full(s), k0 + j * BOXK, 0);
            }
          } else if (MODE == MODE_DGRAD) {   // tile = 128 k; stage = 64 n: 4 W boxes [64 n][32 k] + 2 dz boxes [Mb][32 n]
            const int k0 = tile * 128, n0 = ks * 64;
            for (int j = 0; j < 4; ++j) tma_load_2d(a + j * 8192, &mapW, full(s), k0 + j * BOXK, n0);
            for (int j = 0; j < 2; ++j) tma_load_2d(b + j * row_bytes, &mapDZ, full(s), n0 + j * BOXK, 0);
          } else {                           // tile = (256-n block, 128-k block); one stage: 4 x boxes [Mb][32 k] + Mb/32 dz^T boxes [256 n][32 m]
            const int nb = tile % (p.N / 256), kb = tile / (p.N / 256);
            for (int j = 0; j < 4; ++j) tma_load_2d(a + j * row_bytes, &mapX, full(s), kb * 128 + j * BOXK, 0);
            for (int j = 0; j < p.Mb / 32; ++j) tma_load_2d(b + j * 32768, &mapDZ, full(s), j * BOXK, nb * 256);
          }
        }
      }
    }
  } else if (warp == 1) {
    // ---- MMA issuer: warp-uniform loop, one elected lane issues (see elect_one() in tc_common.cuh)
    const uint32_t id = MODE == MODE_FWD ? idesc(128, NW, 0, 0) : idesc(128, NW, 1, 0);
    int s = 0, lt = 0;
    uint32_t ph = 0;
    for (int tile = blockIdx.x; tile < p.tiles; tile += gridDim.x, ++lt) {
      const int ab = lt & 1;
      if (lt >= 2) mbar_wait_spin(acc_empty(ab), ((lt >> 1) - 1) & 1);
      tc_fence_after();
      const uint32_t d = tmem + ab * acc_stride;
      for (int ks = 0; ks < p.ksteps; ++ks) {
        mbar_wait_spin(full(s), ph);
        tc_fence_after();
        const uint32_t a = base + s * p.stage_bytes, b = a + p.a_bytes;
        if (elect_one()) {
          if (MODE == MODE_FWD) {
#pragma unroll
            for (int i = 0; i < 8; ++i)      // 64 k = 2 boxes x 4 steps of 8 (32 bytes along the swizzled row)
              umma_tf32(d, smem_desc(a + (i >> 2) * 16384 + (i & 3) * 32, 1024, LAYOUT_SW128), smem_desc(b + (i >> 2) * row_bytes + (i & 3) * 32, 1024, LAYOUT_SW128), id,
                        (ks > 0 || i > 0) ? 1u : 0u);
          } else if (MODE == MODE_DGRAD) {
#pragma unroll
            for (int i = 0; i < 8; ++i)      // 64 n = 8 steps of 8 rows of the W boxes (MN-major A: 32-k groups 8 KB apart); dz K-major
              umma_tf32(d, desc_mn32(a + i * 1024, 8192), smem_desc(b + (i >> 2) * row_bytes + (i & 3) * 32, 1024, LAYOUT_SW128), id, (ks > 0 || i > 0) ? 1u : 0u);
          } else {
            for (int i = 0; i < p.Mb / 8; ++i)   // K = batch rows, 8 per step: x MN-major (32-k groups one box apart), dz^T K-major
              umma_tf32(d, desc_mn32(a + i * 1024, row_bytes), smem_desc(b + (i >> 2) * 32768 + (i & 3) * 32, 1024, LAYOUT_SW128), id, i > 0 ? 1u : 0u);
          }
          tc_commit(empty(s));
        }
        __syncwarp();
        if (++s == p.stages) { s = 0; ph ^= 1; }
      }
      if (elect_one()) tc_commit(acc_full(ab));
      __syncwarp();
    }
  } else if (warp >= 4) {
    // ---- epilogue: thread = accumulator row (TMEM lane)
    const int q = warp & 3;
    int lt = 0;
    for (int tile = blockIdx.x; tile < p.tiles; tile += gridDim.x, ++lt) {
      const int ab = lt & 1;
      mbar_wait(acc_full(ab), (lt >> 1) & 1);
      tc_fence_after();
      const uint32_t src = tmem + ab * acc_stride + ((uint32_t)(q * 32) << 16);
      const int row = q * 32 + lane;
      for (int c = 0; c < NW; c += 32) {
        float v[32];
        tmem_ld32(src + c, v);
        if (MODE == MODE_FWD) {              // partial[split][n][m]: the thread's row is contiguous
          const int nb = tile / p.ksplit, sp = tile % p.ksplit;
          float* dst = p.out + ((size_t)sp * p.N + nb * 128 + row) * p.Mb + c;
#pragma unroll
          for (int e = 0; e < 32; e += 4)
            if (c + e < p.Mb) *reinterpret_cast<float4*>(dst + e) = make_float4(v[e], v[e + 1], v[e + 2], v[e + 3]);
        } else if (MODE == MODE_DGRAD) {     // dx[m][k0 + row]: a warp writes 32 consecutive k per m
          float* dst = p.out + (size_t)tile * 128 + row;
#pragma unroll
          for (int e = 0; e < 32; ++e)
            if (c + e < p.Mb) dst[(size_t)(c + e) * p.K] = v[e];
        } else {                             // dW[n0 + c + e][k0 + row]: a warp writes 32 consecutive k per n
          const int nb = tile % (p.N / 256), kb = tile / (p.N / 256);
          float* dst = p.out + (size_t)(nb * 256 + c) * p.K + (size_t)kb * 128 + row;
#pragma unroll
          for (int e = 0; e < 32; ++e) dst[(size_t)e * p.K] = v[e];
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(acc_empty(ab));
    }
  }
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(tmem_cols) : "memory");
  }
}

// y[m][n] = act(sum_s partial[s][n][m] + bias[n])  (fixed order: deterministic)
__global__ void linear_fwd_reduce_kernel(const float* __restrict__ partial, int ksplit, int N, int Mb, const float* __restrict__ bias, int act, float slope,
                                         float* __restrict__ y) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= N * Mb) return;
  const int n = idx / Mb, m = idx - n * Mb;
  float acc = 0.f;
  for (int s = 0; s < ksplit; ++s) acc += partial[((size_t)s * N + n) * Mb + m];
  y[(size_t)m * N + n] = apply_act(acc + (bias ? bias[n] : 0.f), act, slope);
}

static int make_map(CUtensorMap* m, const void* ptr, uint64_t rows, uint64_t cols, uint32_t box_rows, bool mn32 = false) {
  EncodeTiledFn enc = get_encode();
  if (!enc) { set_error("linear_tc: cuTensorMapEncodeTiled unavailable"); return GDN_ECUDA; }
  cuuint64_t gdim[2] = {cols, rows};
  cuuint64_t gstride[1] = {cols * 4};
  cuuint32_t box[2] = {(cuuint32_t)BOXK, box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<void*>(ptr), gdim, gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   mn32 ? CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B : CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { set_error("linear_tc: cuTensorMapEncodeTiled(rows=%llu cols=%llu box_rows=%u) failed (%d)", (unsigned long long)rows, (unsigned long long)cols, box_rows, (int)r); return GDN_ECUDA; }
  return GDN_OK;
}
static size_t smem_bytes(const Params& p) { return (size_t)p.stages * p.stage_bytes + 4 * 4096 + 8 * (2 * 8 + 6) + 1024; }
}  // namespace lintc
}  // namespace gdn

using namespace gdn;
using namespace gdn::lintc;

extern "C" int gdn_linear_tc_init(void) {
  GDN_CHECK_CUDA(cudaFuncSetAttribute(linear_tc_kernel<MODE_FWD>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_LIMIT));
  GDN_CHECK_CUDA(cudaFuncSetAttribute(linear_tc_kernel<MODE_DGRAD>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_LIMIT));
  GDN_CHECK_CUDA(cudaFuncSetAttribute(linear_tc_kernel<MODE_WGRAD>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_LIMIT));
  return GDN_OK;
}

// shapes the kernels take: batch rows multiple of 16 and <= 64, out features multiple of 128, in features multiple of 256
extern "C" int gdn_linear_tc_supported(int Mb, int N, int K) { return Mb >= 16 && Mb <= 64 && Mb % 16 == 0 && N % 128 == 0 && K % 256 == 0 && K >= 4096; }

static int fwd_ksplit(int N, int K) {
  int s = kNumSMs / (N / 128);
  if (s < 1) s = 1;
  while (s > 1 && (K / 64) % s != 0) --s;       // equal K ranges
  return s;
}
extern "C" size_t gdn_linear_tc_fwd_ws_bytes(int Mb, int N, int K) { return (size_t)fwd_ksplit(N, K) * N * Mb * sizeof(float); }

extern "C" int gdn_linear_tc_fwd(const float* x, const float* w, const float* bias, float* y, int Mb, int N, int K, int act, float slope, float* ws, size_t ws_bytes,
                                 gdn_stream_t s) {
  GDN_CHECK_ARG(x && w && y && ws && gdn_linear_tc_supported(Mb, N, K));
  if (ws_bytes < gdn_linear_tc_fwd_ws_bytes(Mb, N, K)) { set_error("gdn_linear_tc_fwd: workspace too small"); return GDN_EWORKSPACE; }
  Params p = {};
  p.Mb = Mb; p.N = N; p.K = K; p.ksplit = fwd_ksplit(N, K);
  p.a_bytes = 2 * 16384; p.stage_bytes = p.a_bytes + 2 * Mb * 128;
  p.stages = (SMEM_LIMIT - 20 * 1024) / p.stage_bytes; if (p.stages > 8) p.stages = 8;
  p.tiles = (N / 128) * p.ksplit; p.ksteps = K / 64 / p.ksplit; p.out = ws;
  CUtensorMap mw, mx;
  int rc;
  if ((rc = make_map(&mw, w, N, K, 128)) != GDN_OK) return rc;
  if ((rc = make_map(&mx, x, Mb, K, Mb)) != GDN_OK) return rc;
  cudaStream_t st = as_stream(s);
  linear_tc_kernel<MODE_FWD><<<p.tiles < kNumSMs ? p.tiles : kNumSMs, NT, smem_bytes(p), st>>>(mw, mx, mx, p);
  GDN_CHECK_LAUNCH();
  linear_fwd_reduce_kernel<<<(unsigned)cdiv((long long)N * Mb, 256), 256, 0, st>>>(ws, p.ksplit, N, Mb, bias, act, slope, y);
  GDN_CHECK_LAUNCH();
  return GDN_OK;
}

extern "C" int gdn_linear_tc_dgrad(const float* dz, const float* w, float* dx, int Mb, int N, int K, gdn_stream_t s) {
  GDN_CHECK_ARG(dz && w && dx && gdn_linear_tc_supported(Mb, N, K));
  Params p = {};
  p.Mb = Mb; p.N = N; p.K = K;
  p.a_bytes = 4 * 8192; p.stage_bytes = p.a_bytes + 2 * Mb * 128;
  p.stages = (SMEM_LIMIT - 20 * 1024) / p.stage_bytes; if (p.stages > 8) p.stages = 8;
  p.tiles = K / 128; p.ksteps = N / 64; p.out = dx;
  CUtensorMap mw, mz;
  int rc;
  if ((rc = make_map(&mw, w, N, K, 64, true)) != GDN_OK) return rc;
  if ((rc = make_map(&mz, dz, Mb, N, Mb)) != GDN_OK) return rc;
  linear_tc_kernel<MODE_DGRAD><<<p.tiles < kNumSMs ? p.tiles : kNumSMs, NT, smem_bytes(p), as_stream(s)>>>(mw, mz, mz, p);
  GDN_CHECK_LAUNCH();
  return GDN_OK;
}

extern "C" size_t gdn_linear_tc_wgrad_ws_bytes(int Mb, int N, int K) { (void)K; return (size_t)N * Mb * sizeof(float); }

__global__ void transpose_small_kernel(const float* __restrict__ in, float* __restrict__ out, int rows, int cols) {   // out[c][r] = in[r][c]
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= rows * cols) return;
  const int c = idx / rows, r = idx - c * rows;
  out[idx] = in[(size_t)r * cols + c];
}

extern "C" int gdn_linear_tc_wgrad(const float* dz, const float* x, float* dw, int Mb, int N, int K, float* ws, size_t ws_bytes, gdn_stream_t s) {
  GDN_CHECK_ARG(dz && x && dw && ws && gdn_linear_tc_supported(Mb, N, K) && Mb % 32 == 0 && N % 256 == 0 && ((uintptr_t)ws & 15) == 0);
  if (ws_bytes < gdn_linear_tc_wgrad_ws_bytes(Mb, N, K)) { set_error("gdn_linear_tc_wgrad: workspace too small"); return GDN_EWORKSPACE; }
  cudaStream_t st = as_stream(s);
  transpose_small_kernel<<<(unsigned)cdiv((long long)N * Mb, 256), 256, 0, st>>>(dz, ws, Mb, N);     // dz^T [N][Mb]
  GDN_CHECK_LAUNCH();
  Params p = {};
  p.Mb = Mb; p.N = N; p.K = K;
  p.a_bytes = 4 * Mb * 128; p.stage_bytes = p.a_bytes + (Mb / 32) * 32768;
  p.stages = (SMEM_LIMIT - 20 * 1024) / p.stage_bytes; if (p.stages > 8) p.stages = 8;
  p.tiles = (N / 256) * (K / 128); p.ksteps = 1; p.out = dw;
  CUtensorMap mx, mz;
  int rc;
  if ((rc = make_map(&mx, x, Mb, K, Mb, true)) != GDN_OK) return rc;
  if ((rc = make_map(&mz, ws, N, Mb, 256)) != GDN_OK) return rc;
  linear_tc_kernel<MODE_WGRAD><<<p.tiles < kNumSMs ? p.tiles : kNumSMs, NT, smem_bytes(p), st>>>(mx, mx, mz, p);
  GDN_CHECK_LAUNCH();
  return GDN_OK;
}
