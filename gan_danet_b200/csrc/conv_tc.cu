// Implicit-GEMM convolution on the 5th-generation tensor cores (sm_100a):
//   TMA (cp.async.bulk.tensor, 4-D tiled maps over the NHWC activation: the box is a Ht x Wt patch of pixels x 64 channels,
//   the zero padding of the convolution is the TMA out-of-bounds fill, the stride of a strided convolution is the TMA
//   element stride) -> 128-byte-swizzled shared memory -> tcgen05.mma kind::f16 (bf16 operands, fp32 accumulate in TMEM)
//   -> tcgen05.ld epilogue with fused bias / activation / residual.
// Replaces nn.Conv2d at /root/reference/models/generator.py:34,63,108-110,148,188,214,218,222, discriminator.py:63-65 and the
// VGG19 convolutions of losses.py:58 in the `bf16` / `bf16x3` precision modes; the fp32 CUDA-core engine (igemm_simt.cu)
// stays the `fp32` mode.
//
// One K-iteration = one filter tap x one 64-channel chunk:   D[128 px, n_tile] += A[128 px, 64 ch] * W[n_tile, 64 ch]^T
//   forward           : A = x at (ho*stride - pad + kh, wo*stride - pad + kw),            W = w[:, :, kh, kw]
//   data gradient s=1 : A = dy at (h + pad - kh, w + pad - kw),                            W = w[:, :, kh, kw]^T
//   data gradient s=2 : the four output-parity classes (h%2, w%2) are four launches, each a dense stride-1 problem over
//                       the dy grid with its own subset of taps, scattered to the strided output pixels.
// Precision `bf16x3` (bf16 hi+lo split): every operand is hi = bf16(v), lo = bf16(v - hi) and the K loop runs
// the three products hi*hi + lo*hi + hi*lo into the same fp32 accumulator (~16 mantissa bits; 3x the MMA work).
//
// Weight gradient (second kernel): out[co, tap, ci] = sum_px dy[px, co] * x[px + tap, ci] is a GEMM whose K dimension is the
// pixel index, i.e. both operands are "MN-major" in shared memory ([pixel][channel] tiles exactly as TMA delivers them);
// tcgen05 takes them through the a_major/b_major bits of the instruction descriptor, so no transpose pass exists.
// The pixel range is split over CTAs; partial sums go to a workspace and a deterministic reduction writes the OIHW gradient.
#include <cuda_bf16.h>
#include <stdlib.h>
#include <type_traits>
#include "tc_common.cuh"

namespace gdn {
namespace convtc {
using namespace gdn::tc;

constexpr int BM = 128;             // pixels per tile (UMMA M)
constexpr int BK = 64;              // channels per K-iteration: 64 bf16 = one 128-byte swizzled row
constexpr int A_BYTES = BM * BK * 2;  // 16 KB
constexpr int NTHREADS = 256;          // weight-gradient kernel
constexpr int FWD_THREADS = 384;       // forward kernel: 4 control warps + 8 epilogue warps
constexpr int EPI_WARPS = 8;
constexpr int MAX_STAGES = 6;
constexpr int SMEM_LIMIT = 227 * 1024;

// kind::f16 instruction descriptor for bf16 operands: D fp32 (bit 4), A/B format bf16 (1 at [7,10) and [10,13)),
// a_major bit 15 / b_major bit 16 (1 = MN-major), N>>3 at [17,23), M>>4 at [24,29)
__host__ __device__ constexpr uint32_t idesc_bf16(int M, int N, int a_mn, int b_mn) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)a_mn << 15) | ((uint32_t)b_mn << 16) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
// shared-memory matrix descriptor with explicit leading-dimension byte offset (needed by MN-major operands wider than 64)
__device__ __forceinline__ uint64_t smem_desc_lbo(uint32_t addr, uint32_t lbo_bytes, uint32_t sbo_bytes, uint32_t layout_type) {
  return (uint64_t)((addr >> 4) & 0x3FFF) | ((uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16) | ((uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32) | ((uint64_t)1 << 46) |
         ((uint64_t)layout_type << 61);
}
__device__ __forceinline__ void tma_load_3d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1, int c2) {
  asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
               ::"r"(dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(bar), "r"(c0), "r"(c1), "r"(c2) : "memory");
}

struct TapList { int n; int dh[9], dw[9], widx[9]; };

struct FwdParams {
  float* y; int y_pitch; const float* bias; const float* res; int res_pitch; const float* alpha_ptr;
  __nv_bfloat16* y16; int y16_pitch;   // optional bf16 copy of the output = the next convolution's packed operand (y may then be NULL)
  int imgs_per_group, taps_total;   // per-sample weights (CAM): weight tap coordinate = tap + (img / imgs_per_group) * taps_total
  int B, Ho, Wo, Cout;        // output tensor (pixels are written at (i*os + oh0, j*os + ow0))
  int Hc, Wc, os, oh0, ow0;   // extent of the tile grid and the output scatter
  int Wt, Ht, tiles_w, tiles_h;
  int cs;                     // input coordinate = tile coordinate * cs + tap offset (cs = conv stride = TMA element stride)
  int kchunks, n_tile, n_tiles, total_tiles, nsplit, stages, tmem_cols, wt_shift;
  int act; float slope; int vec4; int halo_bo, kgroup;
  int w_resident;             // HALO: all weight tiles of the layer stay in shared memory for the CTA's lifetime (loaded once)
  int col_k;                  // COL mode: channels of the narrow operand (<= 24)
  TapList taps;
};

// Persistent kernel: each CTA walks tiles blockIdx.x, blockIdx.x + gridDim.x, ...  (tile = pixel tile x output-channel tile).
//   warp 0: TMA producer (smem ring runs across tile boundaries)      warp 1: tcgen05.mma issuer      warp 2: TMEM allocator
//   warps 4-11: epilogue.  Two accumulator buffers in TMEM: the epilogue of tile i overlaps the main loop of tile i+1.
// Epilogue: tcgen05.ld gives every thread 32 consecutive channels of ONE pixel; the warp transposes its 32x32 block through a
// swizzled shared-memory stage so that each store instruction writes 4 pixels x 128 contiguous bytes (full sectors).
constexpr int EPI_STAGE_BYTES = 32 * 128;      // per epilogue warp: [32 pixels][32 channels] fp32

// HALO variant (stride-1 3x3 problems with narrow output tiles, which are L2-bound when every tap re-reads its own A tile): the
// tile is one image row of 128 pixels; per 64-channel chunk ONE TMA box of 3 rows x 130 pixels is loaded and the 9 taps are 9
// shifted 128-row windows of it, addressed by moving the start address of the shared-memory descriptor by whole 128-byte rows
// (measured on B200: tcgen05 takes the 128B-swizzle phase from the absolute shared-memory address, exactly as TMA wrote it, so a
// start address that is not 1024-byte aligned needs NO base-offset correction).  A traffic drops 9x -> 3.05x.
constexpr int HALO_W = BM + 2, HALO_ROWS = 3;
constexpr int HALO_TX = HALO_ROWS * HALO_W * 128;          // 49920 bytes per box
constexpr int HALO_SLOT = 50 * 1024;                       // slot pitch (1024-aligned)
constexpr int HALO_SLOTS = 2;
__device__ __forceinline__ uint64_t smem_desc_bo(uint32_t addr, uint32_t sbo_bytes, uint32_t layout_type, uint32_t use_bo) {
  return (uint64_t)((addr >> 4) & 0x3FFF) | ((uint64_t)1 << 16) | ((uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32) | ((uint64_t)1 << 46) |
         ((uint64_t)(use_bo ? ((addr >> 7) & 7) : 0) << 49) | ((uint64_t)layout_type << 61);
}

// [128 px + 2][8 ch] boxes of a narrow (<= 24 channel) operand, no swizzle: tcgen05's canonical "interleave" layout (8 channels contiguous, pixel rows 16 B
// apart).  Used as the MN-major A operand of conv_tc_wgrad_col_kernel and as the K-major A operand of the COL mode below.
constexpr int WR_BOX_BYTES = (BM + 2) * 16;                   // 2080 bytes per box ...
constexpr int WR_SLOT_BYTES = 17 * 128;                       // ... in slots of 2176: TMA wants 128-byte aligned shared-memory destinations
constexpr int WR_COL_BYTES = 16 * WR_SLOT_BYTES;              // 34 KB: 16 slots = the 128 accumulator rows of one product (slots >= 3 * G stay zero)

// COL mode (MODE 2, round 2): data gradient of a 3x3 stride-1 "same" convolution whose OUTPUT side is narrow (the DenseNet growth convolutions,
// generator.py:34: dz has 24 channels).  With one K-iteration per (tap, 64-channel chunk) 24 of 64 K columns did work and nine 18 KB weight tiles were
// streamed per pixel tile (0.11-0.14 ms per layer, neither MMA- nor HBM-bound).  Here the reduction runs over (kw, kh, co) = 3 x 80 columns as ONE
// chain of 15 K steps: A = the im2col'd dz tile (3 * G boxes of 130 pixels, the kw shift is the start address; K-major "interleave" layout),
// B = the packed weights [tap][ci][co], resident in shared memory for the CTA's lifetime as 8-column groups ([n_tile rows][16 B], K-major interleave).
constexpr int COL_KGROUPS = 10;                               // K groups (of 8) per kw: 3 * G <= 9 real + zero padding to a whole number of K = 16 steps
template <int MODE>
__global__ void __launch_bounds__(FWD_THREADS, 1)
conv_tc_fwd_kernel(const __grid_constant__ CUtensorMap mapXhi, const __grid_constant__ CUtensorMap mapXlo,
                   const __grid_constant__ CUtensorMap mapWhi, const __grid_constant__ CUtensorMap mapWlo, const FwdParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* sm = smem_raw + (base - smem_u32(smem_raw));
  constexpr bool HALO = MODE == 1, COL = MODE == 2;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int b_bytes = p.n_tile * 128;
  // !HALO: ring of (A tile | W tile) stages.  HALO: ring of HALO_SLOTS halo boxes, then a ring of p.stages W tiles (full/empty barriers).
  // COL: two im2col'd operand tiles, then the resident weight groups.
  const int sub_bytes = A_BYTES + b_bytes;
  const int stage_bytes = p.kgroup * (HALO ? b_bytes : sub_bytes);
  const uint32_t ring_off = HALO ? HALO_SLOTS * HALO_SLOT : (COL ? 2 * WR_COL_BYTES : 0);
  const int col_wstride = p.n_tile * 16;                   // COL: bytes per weight K group ([n_tile rows][8 columns])
  const uint32_t epi_off = ring_off + (COL ? 3 * COL_KGROUPS * col_wstride : p.stages * stage_bytes);
  const uint32_t bar_base = base + epi_off + EPI_WARPS * EPI_STAGE_BYTES;
  auto full = [&](int s) { return bar_base + 8 * s; };
  auto empty = [&](int s) { return bar_base + 8 * (MAX_STAGES + s); };
  auto acc_full = [&](int b) { return bar_base + 8 * (2 * MAX_STAGES + b); };
  auto acc_empty = [&](int b) { return bar_base + 8 * (2 * MAX_STAGES + 2 + b); };
  auto afull = [&](int b) { return bar_base + 8 * (2 * MAX_STAGES + 4 + b); };
  auto aempty = [&](int b) { return bar_base + 8 * (2 * MAX_STAGES + 6 + b); };
  const uint32_t tmem_slot_addr = bar_base + 8 * (2 * MAX_STAGES + 8);
  volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(sm + (tmem_slot_addr - base));
  const int KI = p.taps.n * p.nsplit * p.kchunks;          // K-iterations per tile
  const int acc_stride = p.tmem_cols >> 1;                 // column distance of the two accumulator buffers

  if (COL) {      // zero padding groups of both operands (never written by TMA)
    uint4* z = reinterpret_cast<uint4*>(sm);
    for (int i = threadIdx.x; i < (int)(epi_off >> 4); i += FWD_THREADS) z[i] = make_uint4(0u, 0u, 0u, 0u);
    fence_async_smem();
  }
  if (threadIdx.x == 0) {
    for (int s = 0; s < p.stages; ++s) { mbar_init(full(s), 1); mbar_init(empty(s), 1); }
    for (int b = 0; b < 2; ++b) { mbar_init(acc_full(b), 1); mbar_init(acc_empty(b), EPI_WARPS); mbar_init(afull(b), 1); mbar_init(aempty(b), 1); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot_addr), "r"((uint32_t)p.tmem_cols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;

  if (warp == 0) {
    if (lane == 0) {   // ---- TMA producer
      int it = 0, na = 0;
      if (COL) {
        const int G = (p.col_k + 7) >> 3;
        mbar_expect_tx(full(0), 9 * G * col_wstride);
        for (int kw = 0; kw < 3; ++kw)
          for (int kh = 0; kh < 3; ++kh)
            for (int g = 0; g < G; ++g)
              tma_load_3d(base + ring_off + (kw * COL_KGROUPS + kh * G + g) * col_wstride, &mapWhi, full(0), g * 8, 0, kh * 3 + kw);
        for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x, ++na) {
          int t = tile;
          const int tw = t % p.tiles_w; t /= p.tiles_w;
          const int th = t % p.tiles_h; t /= p.tiles_h;
          const int sa = na & 1;
          if (na >= 2) mbar_wait(aempty(sa), ((na >> 1) - 1) & 1);
          mbar_expect_tx(afull(sa), 3 * G * WR_BOX_BYTES);
          for (int kh = 0; kh < 3; ++kh)
            for (int g = 0; g < G; ++g)      // pixels tw*128 - 1 .. tw*128 + 128 of image row th - (kh - 1): the kw shift is the window's start address
              tma_load_4d(base + sa * WR_COL_BYTES + (kh * G + g) * WR_SLOT_BYTES, &mapXhi, afull(sa), g * 8, tw * BM - 1, th - (kh - 1), t);
        }
      } else {
      if (HALO && p.w_resident) {
        // small layers (e.g. 64 -> 64: 9 x 8 KB): the weights are the same for every tile of this persistent CTA -- load them once instead
        // of once per tile (ncu on the 64 -> 64 layer at 256x512: L2 -> SM traffic 122 KB per tile, 72 KB of it weights; tensor pipe 28 %)
        const int nk = p.kchunks * p.nsplit;
        mbar_expect_tx(full(0), nk * p.taps.n * b_bytes);
        for (int kc = 0; kc < p.kchunks; ++kc)
          for (int comp = 0; comp < p.nsplit; ++comp)
            for (int tp = 0; tp < p.taps.n; ++tp)
              tma_load_3d(base + ring_off + ((kc * p.nsplit + comp) * p.taps.n + tp) * b_bytes, comp == 2 ? &mapWlo : &mapWhi, full(0), kc * BK, 0, p.taps.widx[tp]);
      }
      for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
        int t = tile / p.n_tiles;
        const int n0 = (tile % p.n_tiles) * p.n_tile;
        const int tw = t % p.tiles_w; t /= p.tiles_w;
        const int th = t % p.tiles_h; t /= p.tiles_h;
        const int img = t;
        const int wgrp = (img / p.imgs_per_group) * p.taps_total;
        // one mbarrier round trip per GROUP of p.kgroup K-iterations: with narrow output tiles an MMA K-iteration is only
        // 2*n_tile cycles, far less than a producer/issuer hand-shake
        if (HALO) {
          for (int kc = 0; kc < p.kchunks; ++kc) {
            for (int comp = 0; comp < p.nsplit; ++comp, ++na) {
              const CUtensorMap* mx = comp == 1 ? &mapXlo : &mapXhi;
              const CUtensorMap* mw = comp == 2 ? &mapWlo : &mapWhi;
              const int sa = na % HALO_SLOTS;
              if (na >= HALO_SLOTS) mbar_wait(aempty(sa), ((na / HALO_SLOTS) - 1) & 1);
              mbar_expect_tx(afull(sa), HALO_TX);
              tma_load_4d(base + sa * HALO_SLOT, mx, afull(sa), kc * BK, tw * BM - 1, th - 1, img);
              if (p.w_resident) continue;
              for (int tp0 = 0; tp0 < p.taps.n; tp0 += p.kgroup, ++it) {
                const int s = it % p.stages;
                if (it >= p.stages) mbar_wait(empty(s), ((it / p.stages) - 1) & 1);
                const int cnt = min(p.kgroup, p.taps.n - tp0);
                mbar_expect_tx(full(s), cnt * b_bytes);
                for (int g = 0; g < cnt; ++g)
                  tma_load_3d(base + ring_off + s * stage_bytes + g * b_bytes, mw, full(s), kc * BK, n0, p.taps.widx[tp0 + g] + wgrp);
              }
            }
          }
          continue;
        }
        if (p.kgroup == 1) {                               // wide tiles: one K-iteration per stage, plain nested loops
          for (int tp = 0; tp < p.taps.n; ++tp) {
            const int cw = tw * p.Wt * p.cs + p.taps.dw[tp], ch = th * p.Ht * p.cs + p.taps.dh[tp];
            for (int comp = 0; comp < p.nsplit; ++comp) {
              const CUtensorMap* mx = comp == 1 ? &mapXlo : &mapXhi;
              const CUtensorMap* mw = comp == 2 ? &mapWlo : &mapWhi;
              for (int kc = 0; kc < p.kchunks; ++kc, ++it) {
                const int s = it % p.stages;
                if (it >= p.stages) mbar_wait(empty(s), ((it / p.stages) - 1) & 1);
                mbar_expect_tx(full(s), sub_bytes);
                tma_load_4d(base + s * sub_bytes, mx, full(s), kc * BK, cw, ch, img);
                tma_load_3d(base + s * sub_bytes + A_BYTES, mw, full(s), kc * BK, n0, p.taps.widx[tp] + wgrp);
              }
            }
          }
          continue;
        }
        int tp = 0, comp = 0, kc = 0;                      // K-iteration order: tap, precision component, channel chunk
        for (int k0 = 0; k0 < KI; k0 += p.kgroup, ++it) {
          const int s = it % p.stages;
          if (it >= p.stages) mbar_wait(empty(s), ((it / p.stages) - 1) & 1);
          const int cnt = min(p.kgroup, KI - k0);
          mbar_expect_tx(full(s), cnt * sub_bytes);
          for (int g = 0; g < cnt; ++g) {
            const CUtensorMap* mx = comp == 1 ? &mapXlo : &mapXhi;
            const CUtensorMap* mw = comp == 2 ? &mapWlo : &mapWhi;
            const uint32_t dst = base + s * stage_bytes + g * sub_bytes;
            tma_load_4d(dst, mx, full(s), kc * BK, tw * p.Wt * p.cs + p.taps.dw[tp], th * p.Ht * p.cs + p.taps.dh[tp], img);
            tma_load_3d(dst + A_BYTES, mw, full(s), kc * BK, n0, p.taps.widx[tp] + wgrp);
            if (++kc == p.kchunks) { kc = 0; if (++comp == p.nsplit) { comp = 0; ++tp; } }
          }
        }
      }
      }
    }
  } else if (warp == 1) {
    // ---- MMA issuer.  The whole warp runs the (uniform) loops and waits; one elected lane issues the tcgen05 instructions, whose
    // descriptors are uniform values built from a base descriptor plus 16-byte-unit offsets (see elect_one() in tc_common.cuh:
    // an `if (lane == 0)` issue region costs ~90 cycles per tcgen05.mma, more than a 128 x n_tile x 16 MMA with n_tile <= 128 lasts).
    const uint32_t idesc = idesc_bf16(BM, p.n_tile, 0, 0);
    const uint64_t desc0 = smem_desc(base, 1024, LAYOUT_SW128);       // K-major SWIZZLE_128B tile at `base`; + (byte offset >> 4) moves it
    int s = 0, lt = 0, sa = 0;
    uint32_t ph = 0, pha = 0;
    if (HALO && p.w_resident) mbar_wait_spin(full(0), 0);
    for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x, ++lt) {
      const int ab = lt & 1;
      if (lt >= 2) mbar_wait_spin(acc_empty(ab), ((lt >> 1) - 1) & 1);     // epilogue has drained this accumulator buffer
      tc_fence_after();
      const uint32_t d_tmem = tmem + ab * acc_stride;
      if (COL) {
        if (lt == 0) mbar_wait_spin(full(0), 0);
        mbar_wait_spin(afull(sa), pha);
        tc_fence_after();
        const uint32_t idc = idesc_bf16(BM, p.n_tile, 0, 0);
        // K-major interleave descriptors: LBO = distance of consecutive K groups of 8, SBO = distance of consecutive 8-row groups (128 B)
        const uint64_t a0 = smem_desc_lbo(base + sa * WR_COL_BYTES, WR_SLOT_BYTES, 128, 0);
        const uint64_t b0 = smem_desc_lbo(base + ring_off, (uint32_t)col_wstride, 128, 0);
        if (elect_one()) {
#pragma unroll
          for (int kw = 0; kw < 3; ++kw)
#pragma unroll
            for (int j = 0; j < COL_KGROUPS / 2; ++j)
              umma_f16(d_tmem, a0 + (uint64_t)(2 - kw) + (uint64_t)((2 * j * WR_SLOT_BYTES) >> 4), b0 + (uint64_t)(((kw * COL_KGROUPS + 2 * j) * col_wstride) >> 4), idc,
                       (kw > 0 || j > 0) ? 1u : 0u);
          tc_commit(aempty(sa));
        }
        __syncwarp();
        if (++sa == 2) { sa = 0; pha ^= 1; }
      } else if (HALO && p.w_resident) {
        for (int kk = 0; kk < p.kchunks * p.nsplit; ++kk) {
          mbar_wait_spin(afull(sa), pha);
          tc_fence_after();
          for (int tp = 0; tp < p.taps.n; ++tp) {
            const uint32_t a0 = base + sa * HALO_SLOT + (uint32_t)((p.taps.dh[tp] + 1) * HALO_W + (p.taps.dw[tp] + 1)) * 128u;
            const uint64_t ad = smem_desc_bo(a0, 1024, LAYOUT_SW128, p.halo_bo);
            const uint64_t bd = desc0 + (uint64_t)((ring_off + (kk * p.taps.n + tp) * b_bytes) >> 4);
            if (elect_one()) {
              umma_f16(d_tmem, ad, bd, idesc, (kk > 0 || tp > 0) ? 1u : 0u);
              umma_f16_i<1>(d_tmem, ad + 2, bd + 2, idesc);
              umma_f16_i<1>(d_tmem, ad + 4, bd + 4, idesc);
              umma_f16_i<1>(d_tmem, ad + 6, bd + 6, idesc);
            }
            __syncwarp();
          }
          if (elect_one()) tc_commit(aempty(sa));
          __syncwarp();
          if (++sa == HALO_SLOTS) { sa = 0; pha ^= 1; }
        }
      } else if (HALO) {
        for (int kk = 0; kk < p.kchunks * p.nsplit; ++kk) {
          mbar_wait_spin(afull(sa), pha);
          for (int tp0 = 0; tp0 < p.taps.n; tp0 += p.kgroup) {
            mbar_wait_spin(full(s), ph);
            tc_fence_after();
            const int cnt = min(p.kgroup, p.taps.n - tp0);
            for (int g = 0; g < cnt; ++g) {
              const int tp = tp0 + g;
              // window of tap (dh, dw): halo rows (dh+1)*130 + (dw+1) ... +127, one 128-byte row per pixel
              const uint32_t a0 = base + sa * HALO_SLOT + (uint32_t)((p.taps.dh[tp] + 1) * HALO_W + (p.taps.dw[tp] + 1)) * 128u;
              const uint64_t ad = smem_desc_bo(a0, 1024, LAYOUT_SW128, p.halo_bo);      // + 2 per 32-byte K step stays inside the 128-byte row
              const uint64_t bd = desc0 + (uint64_t)((ring_off + s * stage_bytes + g * b_bytes) >> 4);
              if (elect_one()) {
                umma_f16(d_tmem, ad, bd, idesc, (kk > 0 || tp > 0) ? 1u : 0u);
                umma_f16_i<1>(d_tmem, ad + 2, bd + 2, idesc);
                umma_f16_i<1>(d_tmem, ad + 4, bd + 4, idesc);
                umma_f16_i<1>(d_tmem, ad + 6, bd + 6, idesc);
              }
              __syncwarp();
            }
            if (elect_one()) tc_commit(empty(s));
            __syncwarp();
            if (++s == p.stages) { s = 0; ph ^= 1; }
          }
          if (elect_one()) tc_commit(aempty(sa));
          __syncwarp();
          if (++sa == HALO_SLOTS) { sa = 0; pha ^= 1; }
        }
      } else {
        // a stage holds p.kgroup K-iterations (A tile | W tile) back to back; kgroup == 1 for wide tiles
        for (int k0 = 0; k0 < KI; k0 += p.kgroup) {
          mbar_wait_spin(full(s), ph);
          tc_fence_after();
          const int cnt = min(p.kgroup, KI - k0);
          uint64_t ad = desc0 + (uint64_t)((s * stage_bytes) >> 4);
          for (int g = 0; g < cnt; ++g) {
            const uint64_t bd = ad + (A_BYTES >> 4);
            if (elect_one()) {
              umma_f16(d_tmem, ad, bd, idesc, (k0 + g > 0) ? 1u : 0u);
              umma_f16_i<1>(d_tmem, ad + 2, bd + 2, idesc);
              umma_f16_i<1>(d_tmem, ad + 4, bd + 4, idesc);
              umma_f16_i<1>(d_tmem, ad + 6, bd + 6, idesc);
            }
            __syncwarp();
            ad += (uint64_t)(sub_bytes >> 4);
          }
          if (elect_one()) tc_commit(empty(s));
          __syncwarp();
          if (++s == p.stages) { s = 0; ph ^= 1; }
        }
      }
      if (elect_one()) tc_commit(acc_full(ab));
      __syncwarp();
    }
  } else if (warp >= 4) {
    // ---- epilogue: 8 warps; warp w drains TMEM lanes [32*(w&3), +32) (the hardware's lane quarter of a warp) and the 32-column
    // chunks of parity (w-4)>>2.  Per tile every lane precomputes the output offsets of its 8 store rows once.
    const int q = warp & 3, half = (warp - 4) >> 2;
    float* stg = reinterpret_cast<float*>(sm + epi_off + (warp - 4) * EPI_STAGE_BYTES);
    const float alpha = p.alpha_ptr ? __ldg(p.alpha_ptr) : 1.f;
    const int rsub = lane >> 3, cj = lane & 7;               // store phase: 4 pixel rows per instruction, 8 lanes x 16 B per row
    int lt = 0;
    for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x, ++lt) {
      int t = tile / p.n_tiles;
      const int n0 = (tile % p.n_tiles) * p.n_tile;
      const int tw = t % p.tiles_w; t /= p.tiles_w;
      const int th = t % p.tiles_h; t /= p.tiles_h;
      const int img = t;
      const int ab = lt & 1;
      // Fast path (round 2): the warp's 32 accumulator rows are 32 consecutive pixels of ONE output row that lie inside the image (tile rows of
      // >= 32 pixels, no ragged edge) and all accesses are 128-bit.  Then the addresses of a lane's 8 store rows are base + i8 * step, nothing per
      // row is predicated, and the activation is a compile-time constant of the store loop: ~190 instead of ~900 instructions per 32 x 32 chunk.
      // ncu on the VGG19 64 -> 64 layer (256x512, batch 32): 8 epilogue warps at IPC 1.6 were the kernel's bound (4100 cycles per tile against
      // 1760 of tensor-core work) -- every layer with Cin <= 128 was paced by its epilogue's instruction stream, not by MMA or HBM.
      const int row_w = q * 32;
      const int ti = th * p.Ht + (row_w >> p.wt_shift), tj = tw * p.Wt + (row_w & (p.Wt - 1));
      const bool fast = p.vec4 && p.Wt >= 32 && ti < p.Hc && tj + 31 < p.Wc;
      if (fast) {
        const long long pix0 = ((long long)img * p.Ho + (ti * p.os + p.oh0)) * p.Wo + (tj * p.os + p.ow0) + (long long)rsub * p.os;
        const long long pstep = 4ll * p.os;                      // pixel distance of consecutive store rows of this lane
        mbar_wait(acc_full(ab), (lt >> 1) & 1);
        tc_fence_after();
        const uint32_t src = tmem + ab * acc_stride + ((uint32_t)(q * 32) << 16);
        for (int c = half * 32; c < p.n_tile; c += 64) {
          float v[32];
          tmem_ld32(src + c, v);
#pragma unroll
          for (int j = 0; j < 8; ++j)
            *reinterpret_cast<float4*>(stg + lane * 32 + ((j ^ (lane & 7)) << 2)) = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
          __syncwarp();
          const int n = n0 + c + cj * 4;
          const bool n_ok = n < p.Cout && c + cj * 4 < p.n_tile;
          if (n_ok) {
            const float4 bb = p.bias ? *reinterpret_cast<const float4*>(p.bias + n) : make_float4(0.f, 0.f, 0.f, 0.f);
            float4 rr[8];
            const bool has_res = p.res != nullptr;
            if (has_res) {     // all residual loads first (res may alias y: they must not be serialised behind the stores)
              const float* rp = p.res + pix0 * p.res_pitch + n;
              const long long rstep = pstep * p.res_pitch;
#pragma unroll
              for (int i8 = 0; i8 < 8; ++i8) rr[i8] = *reinterpret_cast<const float4*>(rp + i8 * rstep);
            }
            float* yp = p.y ? p.y + pix0 * p.y_pitch + n : nullptr;
            const long long ystep = pstep * p.y_pitch;
            __nv_bfloat16* hp = p.y16 ? p.y16 + pix0 * p.y16_pitch + n : nullptr;
            const long long hstep = pstep * p.y16_pitch;
            const float* sp = stg + rsub * 32;
            auto rows = [&](auto act_c) {
              constexpr int ACT = decltype(act_c)::value;
#pragma unroll
              for (int i8 = 0; i8 < 8; ++i8) {
                // row r = 4 * i8 + rsub: r & 7 = (4 * (i8 & 1) + rsub) -> the swizzled chunk position alternates between two values
                float4 o = *reinterpret_cast<const float4*>(sp + i8 * 128 + ((cj ^ ((4 * (i8 & 1) + rsub) & 7)) << 2));
                o.x = fmaf(alpha, o.x, bb.x); o.y = fmaf(alpha, o.y, bb.y); o.z = fmaf(alpha, o.z, bb.z); o.w = fmaf(alpha, o.w, bb.w);
                if (ACT == GDN_ACT_RELU) { o.x = o.x > 0.f ? o.x : 0.f; o.y = o.y > 0.f ? o.y : 0.f; o.z = o.z > 0.f ? o.z : 0.f; o.w = o.w > 0.f ? o.w : 0.f; }
                if (ACT == GDN_ACT_LRELU) { o.x = o.x > 0.f ? o.x : o.x * p.slope; o.y = o.y > 0.f ? o.y : o.y * p.slope; o.z = o.z > 0.f ? o.z : o.z * p.slope; o.w = o.w > 0.f ? o.w : o.w * p.slope; }
                if (has_res) { o.x += rr[i8].x; o.y += rr[i8].y; o.z += rr[i8].z; o.w += rr[i8].w; }
                if (yp) *reinterpret_cast<float4*>(yp + i8 * ystep) = o;
                if (hp) {
                  const __nv_bfloat162 h0 = __floats2bfloat162_rn(o.x, o.y), h1 = __floats2bfloat162_rn(o.z, o.w);
                  *reinterpret_cast<uint2*>(hp + i8 * hstep) = make_uint2(*reinterpret_cast<const uint32_t*>(&h0), *reinterpret_cast<const uint32_t*>(&h1));
                }
              }
            };
            if (p.act == GDN_ACT_RELU) rows(std::integral_constant<int, GDN_ACT_RELU>{});
            else if (p.act == GDN_ACT_LRELU) rows(std::integral_constant<int, GDN_ACT_LRELU>{});
            else rows(std::integral_constant<int, GDN_ACT_NONE>{});
          }
          __syncwarp();
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(acc_empty(ab));
        continue;
      }
      long long pixi[8];                                       // output pixel index of this lane's 8 rows (-1: outside the image)
#pragma unroll
      for (int i8 = 0; i8 < 8; ++i8) {
        const int row = q * 32 + 4 * i8 + rsub;
        const int i = th * p.Ht + (row >> p.wt_shift), j = tw * p.Wt + (row & (p.Wt - 1));
        const long long pix = ((long long)img * p.Ho + (i * p.os + p.oh0)) * p.Wo + (j * p.os + p.ow0);
        pixi[i8] = (i < p.Hc && j < p.Wc) ? pix : -1;
      }
      mbar_wait(acc_full(ab), (lt >> 1) & 1);
      tc_fence_after();
      const uint32_t src = tmem + ab * acc_stride + ((uint32_t)(q * 32) << 16);
      for (int c = half * 32; c < p.n_tile; c += 64) {
        float v[32];
        // n_tile is a multiple of 16: the last chunk may be half wide; the columns up to the next multiple of 32 are allocated
        tmem_ld32(src + c, v);
        // transpose through shared memory: row = lane, 16-byte chunk j stored at j ^ (lane & 7)  (conflict-free both ways)
#pragma unroll
        for (int j = 0; j < 8; ++j)
          *reinterpret_cast<float4*>(stg + lane * 32 + ((j ^ (lane & 7)) << 2)) = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
        __syncwarp();
        const int n = n0 + c + cj * 4;
        const bool n_ok = n < p.Cout && c + cj * 4 < p.n_tile;
        float4 bb = make_float4(0.f, 0.f, 0.f, 0.f);
        if (p.bias && n_ok) {
          if (p.vec4) bb = *reinterpret_cast<const float4*>(p.bias + n);
          else { bb.x = __ldg(p.bias + n); if (n + 1 < p.Cout) bb.y = __ldg(p.bias + n + 1); if (n + 2 < p.Cout) bb.z = __ldg(p.bias + n + 2); if (n + 3 < p.Cout) bb.w = __ldg(p.bias + n + 3); }
        }
        if (p.vec4) {
          float4 rr[8];
          if (p.res) {     // all residual loads first (res may alias y: they must not be serialised behind the stores)
#pragma unroll
            for (int i8 = 0; i8 < 8; ++i8)
              rr[i8] = (n_ok && pixi[i8] >= 0) ? *reinterpret_cast<const float4*>(p.res + pixi[i8] * p.res_pitch + n) : make_float4(0.f, 0.f, 0.f, 0.f);
          }
#pragma unroll
          for (int i8 = 0; i8 < 8; ++i8) {
            const int r = 4 * i8 + rsub;
            float4 o = *reinterpret_cast<const float4*>(stg + r * 32 + ((cj ^ (r & 7)) << 2));
            o.x = apply_act(fmaf(alpha, o.x, bb.x), p.act, p.slope); o.y = apply_act(fmaf(alpha, o.y, bb.y), p.act, p.slope);
            o.z = apply_act(fmaf(alpha, o.z, bb.z), p.act, p.slope); o.w = apply_act(fmaf(alpha, o.w, bb.w), p.act, p.slope);
            if (p.res) { o.x += rr[i8].x; o.y += rr[i8].y; o.z += rr[i8].z; o.w += rr[i8].w; }
            if (n_ok && pixi[i8] >= 0) {
              if (p.y) *reinterpret_cast<float4*>(p.y + pixi[i8] * p.y_pitch + n) = o;
              if (p.y16) {
                const __nv_bfloat162 h0 = __floats2bfloat162_rn(o.x, o.y), h1 = __floats2bfloat162_rn(o.z, o.w);
                *reinterpret_cast<uint2*>(p.y16 + pixi[i8] * p.y16_pitch + n) = make_uint2(*reinterpret_cast<const uint32_t*>(&h0), *reinterpret_cast<const uint32_t*>(&h1));
              }
            }
          }
        } else {
#pragma unroll
          for (int i8 = 0; i8 < 8; ++i8) {
            const int r = 4 * i8 + rsub;
            const float4 o4 = *reinterpret_cast<const float4*>(stg + r * 32 + ((cj ^ (r & 7)) << 2));
            if (!(n_ok && pixi[i8] >= 0)) continue;
            const float ov[4] = {o4.x, o4.y, o4.z, o4.w}, bv[4] = {bb.x, bb.y, bb.z, bb.w};
            float* yp = p.y + pixi[i8] * p.y_pitch + n;
            const float* rp = p.res ? p.res + pixi[i8] * p.res_pitch + n : nullptr;
#pragma unroll
            for (int u = 0; u < 4; ++u)
              if (n + u < p.Cout) yp[u] = apply_act(fmaf(alpha, ov[u], bv[u]), p.act, p.slope) + (rp ? rp[u] : 0.f);
          }
        }
        __syncwarp();
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(acc_empty(ab));
    }
  }
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"((uint32_t)p.tmem_cols) : "memory");
  }
}

// ------------------------------------------------------------------------------------------------ weight gradient
constexpr int WG_DY_SLOTS = 2, WG_X_SLOTS = 8;     // WG_X_SLOTS: the most x slots a launch may use (barrier layout); WgParams::x_slots are in use
constexpr int WG_DY_BYTES = 2 * A_BYTES;    // [128 px][64 co] x 2 channel boxes (UMMA M = 128 output channels)

struct WgParams {
  float* direct; int direct_pitch, Cin; const float* scale_ptr;   // grouped mode (one split per sample): out[split][co][ci] = scale * acc, no reduction pass
  float* ws;                   // [splits][Cout][taps][cin_w] fp32 partial sums
  int Cout, cin_w, taps_total;
  int tiles_w, tiles_h, tiles_total, tiles_per_split;
  int Wt, Ht, cs, pad, kw;
  int ci_tile, ci_tiles, co_tiles, tap_groups, taps_per_cta, nsplit, tmem_cols;
  int x_slots;                 // depth of the shifted-operand ring: the kernel is bound by bytes in flight per SM (TMA latency x ring depth), ncu round 2
};

__global__ void __launch_bounds__(NTHREADS)
conv_tc_wgrad_kernel(const __grid_constant__ CUtensorMap mapDYhi, const __grid_constant__ CUtensorMap mapDYlo,
                     const __grid_constant__ CUtensorMap mapXhi, const __grid_constant__ CUtensorMap mapXlo, const WgParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nbox = p.ci_tile / 64;
  const int x_bytes = nbox * A_BYTES;
  const uint32_t dy_base = base, x_base = base + WG_DY_SLOTS * WG_DY_BYTES;
  const int XS = p.x_slots;
  const uint32_t bar_base = x_base + XS * x_bytes;
  auto dyfull = [&](int s) { return bar_base + 8 * s; };
  auto dyempty = [&](int s) { return bar_base + 8 * (WG_DY_SLOTS + s); };
  auto xfull = [&](int s) { return bar_base + 8 * (2 * WG_DY_SLOTS + s); };
  auto xempty = [&](int s) { return bar_base + 8 * (2 * WG_DY_SLOTS + WG_X_SLOTS + s); };
  const uint32_t acc_bar = bar_base + 8 * (2 * WG_DY_SLOTS + 2 * WG_X_SLOTS);
  const uint32_t tmem_slot_addr = acc_bar + 8;
  volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(smem_raw + (tmem_slot_addr - smem_u32(smem_raw)));

  int u = blockIdx.x;
  const int tg = u % p.tap_groups; u /= p.tap_groups;
  const int cit = u % p.ci_tiles; u /= p.ci_tiles;
  const int cot = u;
  const int split = blockIdx.y;
  const int t_begin = split * p.tiles_per_split;
  const int t_end = min(t_begin + p.tiles_per_split, p.tiles_total);
  const int ntiles = t_end - t_begin;          // > 0 by construction of the grid
  const int tap0 = tg * p.taps_per_cta;

  if (threadIdx.x == 0) {
    for (int s = 0; s < WG_DY_SLOTS; ++s) { mbar_init(dyfull(s), 1); mbar_init(dyempty(s), 1); }
    for (int s = 0; s < XS; ++s) { mbar_init(xfull(s), 1); mbar_init(xempty(s), 1); }
    mbar_init(acc_bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot_addr), "r"((uint32_t)p.tmem_cols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;

  if (warp == 0) {
    if (lane == 0) {   // ---- TMA producer for dy tiles: one (pixel tile, precision component) per slot
      int n = 0;
      for (int ti = 0; ti < ntiles; ++ti) {
        int t = t_begin + ti;
        const int tw = t % p.tiles_w; t /= p.tiles_w;
        const int th = t % p.tiles_h; t /= p.tiles_h;
        for (int comp = 0; comp < p.nsplit; ++comp, ++n) {
          const int s = n % WG_DY_SLOTS;
          if (n >= WG_DY_SLOTS) mbar_wait(dyempty(s), ((n / WG_DY_SLOTS) - 1) & 1);
          const CUtensorMap* m = comp == 1 ? &mapDYlo : &mapDYhi;
          mbar_expect_tx(dyfull(s), WG_DY_BYTES);
          tma_load_4d(dy_base + s * WG_DY_BYTES, m, dyfull(s), cot * 128, tw * p.Wt, th * p.Ht, t);
          tma_load_4d(dy_base + s * WG_DY_BYTES + A_BYTES, m, dyfull(s), cot * 128 + 64, tw * p.Wt, th * p.Ht, t);
        }
      }
    }
  } else if (warp == 3) {
    if (lane == 0) {   // ---- TMA producer for the shifted x tiles: one (pixel tile, component, tap) per slot
      int n = 0;
      for (int ti = 0; ti < ntiles; ++ti) {
        int t = t_begin + ti;
        const int tw = t % p.tiles_w; t /= p.tiles_w;
        const int th = t % p.tiles_h; t /= p.tiles_h;
        for (int comp = 0; comp < p.nsplit; ++comp) {
          const CUtensorMap* m = comp == 2 ? &mapXlo : &mapXhi;
          for (int tl = 0; tl < p.taps_per_cta; ++tl, ++n) {
            const int tap = tap0 + tl;
            const int dh = tap / p.kw - p.pad, dw = tap % p.kw - p.pad;
            const int s = n % XS;
            if (n >= XS) mbar_wait(xempty(s), ((n / XS) - 1) & 1);
            mbar_expect_tx(xfull(s), x_bytes);
            for (int bx = 0; bx < nbox; ++bx)
              tma_load_4d(x_base + s * x_bytes + bx * A_BYTES, m, xfull(s), cit * p.ci_tile + bx * 64, tw * p.Wt * p.cs + dw, th * p.Ht * p.cs + dh, t);
          }
        }
      }
    }
  } else if (warp == 1) {
    // ---- MMA issuer (warp-uniform loop, one elected lane issues): acc[tap] (128 co x ci_tile) += dy^T (MN-major A) * x (MN-major B),
    // K = 128 pixels per tile
    const uint32_t idesc = idesc_bf16(128, p.ci_tile, 1, 1);
    const uint64_t desc0 = smem_desc_lbo(base, A_BYTES, 1024, LAYOUT_SW128);   // channel groups of 64 are LBO = 16 KB apart, 8-pixel atoms SBO = 1 KB
    int sd = 0, sx = 0;
    uint32_t phd = 0, phx = 0;
    for (int ti = 0; ti < ntiles; ++ti) {
      for (int comp = 0; comp < p.nsplit; ++comp) {
        mbar_wait_spin(dyfull(sd), phd);
        const uint64_t ad = desc0 + (uint64_t)((sd * WG_DY_BYTES) >> 4);
        for (int tl = 0; tl < p.taps_per_cta; ++tl) {
          mbar_wait_spin(xfull(sx), phx);
          tc_fence_after();
          const uint64_t bd = desc0 + (uint64_t)((WG_DY_SLOTS * WG_DY_BYTES + sx * x_bytes) >> 4);
          const uint32_t d = tmem + tl * p.ci_tile;
          if (elect_one()) {
            umma_f16(d, ad, bd, idesc, (ti > 0 || comp > 0) ? 1u : 0u);
#pragma unroll
            for (int ks = 1; ks < BM / 16; ++ks)     // 16 pixels = two 8-row swizzle atoms = 2 KB
              umma_f16_i<1>(d, ad + ks * 128, bd + ks * 128, idesc);
            tc_commit(xempty(sx));
          }
          __syncwarp();
          if (++sx == XS) { sx = 0; phx ^= 1; }
        }
        if (elect_one()) tc_commit(dyempty(sd));
        __syncwarp();
        if (++sd == WG_DY_SLOTS) { sd = 0; phd ^= 1; }
      }
    }
    if (elect_one()) tc_commit(acc_bar);
    __syncwarp();
  } else if (warp >= 4) {
    // ---- epilogue: thread = output channel (TMEM lane); partial sums to the workspace
    const int q = warp & 3;
    const int co = cot * 128 + q * 32 + lane;
    mbar_wait(acc_bar, 0);
    tc_fence_after();
    const float scale = (p.direct && p.scale_ptr) ? __ldg(p.scale_ptr) : 1.f;
    for (int tl = 0; tl < p.taps_per_cta; ++tl) {
      float* dst = p.direct ? p.direct + ((size_t)split * p.Cout + co) * p.direct_pitch + (size_t)cit * p.ci_tile
                            : p.ws + (((size_t)split * p.Cout + co) * p.taps_total + (tap0 + tl)) * p.cin_w + (size_t)cit * p.ci_tile;
      for (int c = 0; c < p.ci_tile; c += 32) {
        float v[32];
        tmem_ld32(tmem + ((uint32_t)(q * 32) << 16) + tl * p.ci_tile + c, v);
        if (co < p.Cout) {
#pragma unroll
          for (int e = 0; e < 32; e += 4) {
            if (p.direct && cit * p.ci_tile + c + e >= p.Cin) break;      // Cin is a multiple of 4 in direct mode
            *reinterpret_cast<float4*>(dst + c + e) = make_float4(scale * v[e], scale * v[e + 1], scale * v[e + 2], scale * v[e + 3]);
          }
        }
      }
    }
    tc_fence_before();
  }
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"((uint32_t)p.tmem_cols) : "memory");
  }
}


// ------------------------------------------------------------------------------------------------ narrow-output weight gradient (round 2)
// The DenseNet growth convolutions (generator.py:34: C_i -> 24, 3x3, twelve per step) ran their weight gradient at 66-125 TFLOP/s on the kernel
// above: nine shifted x tiles (16-32 KB each) per pixel tile feed nine tiny MMAs, 1400 cycles per tap against ~200 of tensor work, 9x the
// x traffic.  Here the NARROW operand is shifted instead: per 128-pixel tile q the shared-memory tile
//     dycol[q][(tap, co)] = dy[q - (tap - pad)][co]          (216 columns for Cout = 24, zero outside the image)
// is assembled by 27 small TMA boxes ([128 px][8 ch], 16-byte rows, no swizzle -- exactly the canonical MN-major "interleave" layout of
// tcgen05: 8 channels contiguous, pixel rows 16 B apart, 8-pixel groups 128 B apart, channel groups 2 KB apart), x is loaded ONCE, and
//     D[(tap, co)][ci] += dycol^T x                          (M = 256 = two UMMA halves, N = round_up(Cin, 16), K = 128 pixels)
// is one chain of 16 MMAs per tile into TMEM-resident accumulators that live for the CTA's whole lifetime (persistent CTAs, one partial sum
// per CTA, fixed-order reduction => deterministic).  Operand traffic per tile: 54 KB + x instead of 9 x (x + dy).
constexpr int WC_GROUP_BYTES = BM * 16;                       // [128 px][8 ch] bf16
constexpr int WC_GROUPS = 32;                                 // 256 accumulator rows; groups >= 9 * G stay zero
constexpr int WC_COL_BYTES = WC_GROUPS * WC_GROUP_BYTES;      // 64 KB
// ROW variant (tiles of one image row, Wt = 128): the three kw-shifts of a (kh, channel group) are ONE box of 130 pixels read at start addresses
// +0 / +16 / +32 bytes (no swizzle: any 16-byte aligned start is legal), so 3 * G boxes per tile instead of 9 * G (TMA moves these boxes as 16-byte
// rows: the row count, not the byte count, is what they cost).  Accumulator rows are then ordered (kw, kh, co): three M = 128 products per K step.
constexpr int WC_STAGES = 2;
struct WcParams {
  float* ws;                   // [ctas][taps * Cout][N] fp32 partial sums
  int Cout, G, ngroups, nbox, N, tmem_cols;
  int tiles_w, tiles_h, tiles_total, Wt, Ht;
  int lbo, sbo;                // A descriptor strides (bytes): 8-pixel groups / 8-channel groups
};

template <bool ROW>
__global__ void __launch_bounds__(NTHREADS)
conv_tc_wgrad_col_kernel(const __grid_constant__ CUtensorMap mapDY, const __grid_constant__ CUtensorMap mapX, const WcParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* sm = smem_raw + (base - smem_u32(smem_raw));
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  constexpr int COL_BYTES = ROW ? WR_COL_BYTES : WC_COL_BYTES;
  const int stage_bytes = COL_BYTES + p.nbox * A_BYTES;
  const uint32_t bar_base = base + WC_STAGES * stage_bytes;
  auto full = [&](int s) { return bar_base + 8 * s; };
  auto empty = [&](int s) { return bar_base + 8 * (WC_STAGES + s); };
  const uint32_t acc_bar = bar_base + 8 * (2 * WC_STAGES);
  const uint32_t tmem_slot_addr = acc_bar + 8;
  volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(sm + (tmem_slot_addr - base));
  const int ntiles = (p.tiles_total - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;     // tiles blockIdx.x, + gridDim.x, ...
  const int col_tx = ROW ? 3 * p.G * WR_BOX_BYTES : p.ngroups * WC_GROUP_BYTES;                    // bytes TMA writes per stage

  for (int s = 0; s < WC_STAGES; ++s) {      // the groups TMA never writes (and the slot padding) must read as zero
    uint4* z = reinterpret_cast<uint4*>(sm + s * stage_bytes);
    for (int i = threadIdx.x; i < COL_BYTES / 16; i += NTHREADS) z[i] = make_uint4(0u, 0u, 0u, 0u);
  }
  fence_async_smem();
  if (threadIdx.x == 0) {
    for (int s = 0; s < WC_STAGES; ++s) { mbar_init(full(s), 1); mbar_init(empty(s), 1); }
    mbar_init(acc_bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot_addr), "r"((uint32_t)p.tmem_cols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;

  if (warp == 0) {
    if (lane == 0) {   // ---- TMA producer: the shifted dy boxes + the x tile per pixel tile
      for (int n = 0; n < ntiles; ++n) {
        int t = blockIdx.x + n * gridDim.x;
        const int tw = t % p.tiles_w; t /= p.tiles_w;
        const int th = t % p.tiles_h; t /= p.tiles_h;
        const int s = n % WC_STAGES;
        if (n >= WC_STAGES) mbar_wait(empty(s), ((n / WC_STAGES) - 1) & 1);
        const uint32_t st = base + s * stage_bytes;
        mbar_expect_tx(full(s), col_tx + p.nbox * A_BYTES);
        for (int bx = 0; bx < p.nbox; ++bx) tma_load_4d(st + COL_BYTES + bx * A_BYTES, &mapX, full(s), bx * 64, tw * p.Wt, th * p.Ht, t);
        if (ROW) {
          for (int kh = 0; kh < 3; ++kh)
            for (int g = 0; g < p.G; ++g)      // pixels tw*128 - 1 .. tw*128 + 128 of image row th - (kh - 1)
              tma_load_4d(st + (kh * p.G + g) * WR_SLOT_BYTES, &mapDY, full(s), g * 8, tw * BM - 1, th - (kh - 1), t);
        } else {
          for (int tap = 0; tap < 9; ++tap) {
            const int dh = tap / 3 - 1, dw = tap % 3 - 1;
            for (int g = 0; g < p.G; ++g)
              tma_load_4d(st + (tap * p.G + g) * WC_GROUP_BYTES, &mapDY, full(s), g * 8, tw * p.Wt - dw, th * p.Ht - dh, t);
          }
        }
      }
    }
  } else if (warp == 1) {
    // ---- MMA issuer: A = dycol (MN-major, no swizzle), B = x (MN-major, SWIZZLE_128B, 64-channel boxes LBO = 16 KB apart)
    const uint32_t idesc = idesc_bf16(128, p.N, 1, 1);
    const uint64_t a0 = smem_desc_lbo(base, (uint32_t)p.lbo, (uint32_t)p.sbo, 0);
    const uint64_t b0 = smem_desc_lbo(base, A_BYTES, 1024, LAYOUT_SW128);
    for (int n = 0; n < ntiles; ++n) {
      const int s = n % WC_STAGES;
      mbar_wait_spin(full(s), (uint32_t)(n / WC_STAGES) & 1u);
      tc_fence_after();
      const uint64_t ad = a0 + (uint64_t)((s * stage_bytes) >> 4), bd = b0 + (uint64_t)((s * stage_bytes + COL_BYTES) >> 4);
      if (elect_one()) {
#pragma unroll
        for (int ks = 0; ks < BM / 16; ++ks) {       // 16 pixels: 256 B of every dy group, 2 KB of every x box
          const uint32_t acc = (n > 0 || ks > 0) ? 1u : 0u;
          if (ROW) {
            // window kw starts 2 - kw pixels into the 130-pixel box: dycol[q][kw] = dy[q - (kw - 1)] = box pixel (q + 1) - (kw - 1)
            umma_f16(tmem, ad + ks * 16 + 2, bd + ks * 128, idesc, acc);
            umma_f16(tmem + p.N, ad + ks * 16 + 1, bd + ks * 128, idesc, acc);
            umma_f16(tmem + 2 * p.N, ad + ks * 16, bd + ks * 128, idesc, acc);
          } else {
            umma_f16(tmem, ad + ks * 16, bd + ks * 128, idesc, acc);
            umma_f16(tmem + p.N, ad + ((16 * WC_GROUP_BYTES) >> 4) + ks * 16, bd + ks * 128, idesc, acc);
          }
        }
        tc_commit(empty(s));
      }
      __syncwarp();
    }
    if (elect_one()) tc_commit(acc_bar);
    __syncwarp();
  } else if (warp >= 4) {
    // ---- epilogue: thread = accumulator row; this CTA's partial sums to the workspace
    const int q = warp & 3;
    mbar_wait(acc_bar, 0);
    tc_fence_after();
    for (int h = 0; h < (ROW ? 3 : 2); ++h) {
      const int r = q * 32 + lane;
      int tap, co; bool ok;
      if (ROW) { const int i = r >> 3; tap = (i / p.G) * 3 + h; co = (i % p.G) * 8 + (r & 7); ok = i < 3 * p.G && co < p.Cout; }     // accumulator h = kw, row = (kh, co)
      else { const int g = (h * 128 + r) >> 3; tap = g / p.G; co = (g % p.G) * 8 + (r & 7); ok = g < p.ngroups && co < p.Cout; }
      float* dst = p.ws + (((size_t)blockIdx.x * 9 + tap) * p.Cout + co) * p.N;
      for (int c = 0; c < p.N; c += 32) {
        float v[32];
        tmem_ld32(tmem + ((uint32_t)(q * 32) << 16) + h * p.N + c, v);
        if (ok) {
#pragma unroll
          for (int e = 0; e < 32; e += 4)
            if (c + e < p.N) *reinterpret_cast<float4*>(dst + c + e) = make_float4(v[e], v[e + 1], v[e + 2], v[e + 3]);
        }
      }
    }
    tc_fence_before();
  }
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"((uint32_t)p.tmem_cols) : "memory");
  }
}
// out[co][out_c0+ci][tap] (OIHW) (+)= scale * sum_cta ws[cta][tap][co][ci]   -- fixed summation order => deterministic
__global__ void wgrad_col_reduce_kernel(const float* __restrict__ ws, int ctas, int Cout, int Cin, int N, float* __restrict__ out, int out_cin_total, int out_c0,
                                        int accumulate, float scale, const float* __restrict__ scale_ptr) {
  if (scale_ptr) scale *= __ldg(scale_ptr);
  const long long total = (long long)Cout * 9 * Cin;
  const size_t cta_stride = (size_t)9 * Cout * N;
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
    const int ci = (int)(idx % Cin); long long r = idx / Cin;
    const int co = (int)(r % Cout); const int tap = (int)(r / Cout);
    const float* src = ws + ((size_t)tap * Cout + co) * N + ci;
    float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
    int c = 0;
    for (; c + 3 < ctas; c += 4) { a0 += src[c * cta_stride]; a1 += src[(c + 1) * cta_stride]; a2 += src[(c + 2) * cta_stride]; a3 += src[(c + 3) * cta_stride]; }
    for (; c < ctas; ++c) a0 += src[c * cta_stride];
    const float acc = (a0 + a1) + (a2 + a3);
    float* o = out + ((size_t)co * out_cin_total + out_c0 + ci) * 9 + tap;
    *o = accumulate ? *o + acc * scale : acc * scale;
  }
}

// out[co][out_c0+ci][tap] (OIHW) (+)= scale * sum_s ws[s][co][tap][ci]   -- fixed summation order => deterministic
// swapped: the partial sums come from the role-swapped launch (gdn_conv2d_wgrad_tc: narrow Cout), ws[s][ci][taps-1-tap][co] with row width cin_w
__global__ void wgrad_tc_reduce_kernel(const float* __restrict__ ws, int splits, int Cout, int Cin, int taps, int cin_w,
                                       float* __restrict__ out, int out_cin_total, int out_c0, int accumulate, float scale,
                                       const float* __restrict__ scale_ptr, int swapped) {
  if (scale_ptr) scale *= __ldg(scale_ptr);
  const long long total = (long long)Cout * taps * Cin;
  const size_t split_stride = (size_t)(swapped ? Cin : Cout) * taps * cin_w;
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
    const int ci = (int)(idx % Cin); long long r = idx / Cin;
    const int tap = (int)(r % taps); const int co = (int)(r / taps);
    const float* src = swapped ? ws + ((size_t)ci * taps + (taps - 1 - tap)) * cin_w + co : ws + ((size_t)co * taps + tap) * cin_w + ci;
    float acc = 0.f;
    for (int s = 0; s < splits; ++s) acc += src[s * split_stride];
    float* o = out + ((size_t)co * out_cin_total + out_c0 + ci) * taps + tap;
    *o = accumulate ? *o + acc * scale : acc * scale;
  }
}

// ------------------------------------------------------------------------------------------------ operand packing
// fp32 NHWC slice -> bf16 [M][Cp] (Cp = round_up(C, 8), zero padded): hi = bf16(v), lo = bf16(v - hi); v = act(x*scale+shift), or with a
// gate tensor v = x * act'(gate) (activation backward fused into the operand packing of the data/weight gradient: the fp32 dz is never stored)
__global__ void __launch_bounds__(256) pack_act_kernel(const float* __restrict__ x, int pitch, long long M, int C, int Cp, __nv_bfloat16* __restrict__ hi,
                                                       __nv_bfloat16* __restrict__ lo, const float* __restrict__ scale, const float* __restrict__ shift, int act, float slope,
                                                       const float* __restrict__ gate, int gate_pitch, const __nv_bfloat16* __restrict__ gate16) {
  const int groups = Cp >> 3;
  const long long total = M * groups;
  const bool vec = (pitch & 3) == 0 && (reinterpret_cast<uintptr_t>(x) & 15) == 0;
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
    long long m; int c0;
    if (total < (1ll << 31)) { const unsigned i32 = (unsigned)idx, m32 = i32 / (unsigned)groups; m = m32; c0 = (int)(i32 - m32 * (unsigned)groups) * 8; }
    else { m = idx / groups; c0 = (int)(idx % groups) * 8; }
    const float* src = x + (size_t)m * pitch + c0;
    __align__(16) __nv_bfloat16 h[8];
    __align__(16) __nv_bfloat16 l[8];
    float vals[8], gvals[8];
    if (gate16) {            // bf16 gate with pitch gate_pitch (a multiple of 8, 16-byte aligned rows): only its sign / zero-ness matters
      const uint4 gq = __ldg(reinterpret_cast<const uint4*>(gate16 + (size_t)m * gate_pitch + c0));
      const __nv_bfloat16* gh = reinterpret_cast<const __nv_bfloat16*>(&gq);
#pragma unroll
      for (int e = 0; e < 8; ++e) gvals[e] = __bfloat162float(gh[e]);
    } else if (gate) {
      const float* gs = gate + (size_t)m * gate_pitch + c0;
      if ((gate_pitch & 3) == 0 && (reinterpret_cast<uintptr_t>(gate) & 15) == 0 && c0 + 8 <= C) {
        const float4 g0 = __ldg(reinterpret_cast<const float4*>(gs)), g1 = __ldg(reinterpret_cast<const float4*>(gs) + 1);
        gvals[0] = g0.x; gvals[1] = g0.y; gvals[2] = g0.z; gvals[3] = g0.w; gvals[4] = g1.x; gvals[5] = g1.y; gvals[6] = g1.z; gvals[7] = g1.w;
      } else {
#pragma unroll
        for (int e = 0; e < 8; ++e) gvals[e] = c0 + e < C ? __ldg(gs + e) : 0.f;
      }
    }
    if (vec && c0 + 8 <= C) {     // 16-byte aligned rows: two 128-bit loads
      const float4 q0 = __ldg(reinterpret_cast<const float4*>(src)), q1 = __ldg(reinterpret_cast<const float4*>(src) + 1);
      vals[0] = q0.x; vals[1] = q0.y; vals[2] = q0.z; vals[3] = q0.w; vals[4] = q1.x; vals[5] = q1.y; vals[6] = q1.z; vals[7] = q1.w;
    } else {
#pragma unroll
      for (int e = 0; e < 8; ++e) vals[e] = c0 + e < C ? __ldg(src + e) : 0.f;
    }
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      float v = vals[e];
      if (scale && c0 + e < C) v = apply_act(fmaf(v, __ldg(scale + c0 + e), __ldg(shift + c0 + e)), act, slope);
      if ((gate || gate16) && c0 + e < C) v *= act_grad(gvals[e], act, slope);
      h[e] = __float2bfloat16_rn(v);
      l[e] = __float2bfloat16_rn(v - __bfloat162float(h[e]));
    }
    *reinterpret_cast<uint4*>(hi + (size_t)m * Cp + c0) = *reinterpret_cast<const uint4*>(h);
    if (lo) *reinterpret_cast<uint4*>(lo + (size_t)m * Cp + c0) = *reinterpret_cast<const uint4*>(l);
  }
}
// OIHW fp32 -> bf16 [taps][R][Kp]: transposed == 0: R = O, K = I (forward operand); transposed == 1: R = I, K = O (data-gradient operand)
__global__ void __launch_bounds__(256) pack_weight_kernel(const float* __restrict__ w, int O, int I_total, int i_c0, int I, int taps, int transposed,
                                                          __nv_bfloat16* __restrict__ hi, __nv_bfloat16* __restrict__ lo) {
  const int R = transposed ? I : O, K = transposed ? O : I, Kp = (K + 7) & ~7;
  const long long total = (long long)taps * R * Kp;
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
    const int k = (int)(idx % Kp); long long r2 = idx / Kp;
    const int r = (int)(r2 % R); const int t = (int)(r2 / R);
    float v = 0.f;
    if (k < K) {
      const int o = transposed ? k : r, i = transposed ? r : k;
      v = __ldg(w + ((size_t)o * I_total + i_c0 + i) * taps + t);
    }
    const __nv_bfloat16 h = __float2bfloat16_rn(v);
    hi[idx] = h;
    if (lo) lo[idx] = __float2bfloat16_rn(v - __bfloat162float(h));
  }
}

// 4-D bf16 NHWC activation map: dims (Cp, W, H, B), box (64, Wt*cs, Ht*cs, 1) traversed with element stride cs
static int make_act_map(CUtensorMap* m, const void* ptr, int Cp, int W, int H, int B, int Wt, int Ht, int cs) {   // box = Wt x Ht pixels (halo: 130 x 3)
  EncodeTiledFn enc = get_encode();
  if (!enc) { set_error("conv_tc: cuTensorMapEncodeTiled unavailable"); return GDN_ECUDA; }
  cuuint64_t gdim[4] = {(cuuint64_t)Cp, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)B};
  cuuint64_t gstride[3] = {(cuuint64_t)Cp * 2, (cuuint64_t)W * Cp * 2, (cuuint64_t)H * W * Cp * 2};
  cuuint32_t box[4] = {(cuuint32_t)BK, (cuuint32_t)(Wt * cs), (cuuint32_t)(Ht * cs), 1};
  cuuint32_t estr[4] = {1, (cuuint32_t)cs, (cuuint32_t)cs, 1};
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(ptr), gdim, gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { set_error("conv_tc: cuTensorMapEncodeTiled(activation Cp=%d W=%d H=%d B=%d box %dx%d cs=%d) failed (%d)", Cp, W, H, B, Wt, Ht, cs, (int)r); return GDN_ECUDA; }
  return GDN_OK;
}
// the same tensor as [128 px][8 ch] boxes without swizzle (conv_tc_wgrad_col_kernel's shifted dy groups)
static int make_act_map8(CUtensorMap* m, const void* ptr, int Cp, int W, int H, int B, int Wt, int Ht) {
  EncodeTiledFn enc = get_encode();
  if (!enc) { set_error("conv_tc: cuTensorMapEncodeTiled unavailable"); return GDN_ECUDA; }
  cuuint64_t gdim[4] = {(cuuint64_t)Cp, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)B};
  cuuint64_t gstride[3] = {(cuuint64_t)Cp * 2, (cuuint64_t)W * Cp * 2, (cuuint64_t)H * W * Cp * 2};
  cuuint32_t box[4] = {8, (cuuint32_t)Wt, (cuuint32_t)Ht, 1};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(ptr), gdim, gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { set_error("conv_tc: cuTensorMapEncodeTiled(8-channel groups Cp=%d W=%d H=%d B=%d box %dx%d) failed (%d)", Cp, W, H, B, Wt, Ht, (int)r); return GDN_ECUDA; }
  return GDN_OK;
}
// 3-D packed weight map: dims (Kp, R, taps), box (64, n_tile, 1)
static int make_weight_map(CUtensorMap* m, const void* ptr, int Kp, int R, int taps, int n_tile) {
  EncodeTiledFn enc = get_encode();
  if (!enc) { set_error("conv_tc: cuTensorMapEncodeTiled unavailable"); return GDN_ECUDA; }
  cuuint64_t gdim[3] = {(cuuint64_t)Kp, (cuuint64_t)R, (cuuint64_t)taps};
  cuuint64_t gstride[2] = {(cuuint64_t)Kp * 2, (cuuint64_t)R * Kp * 2};
  cuuint32_t box[3] = {(cuuint32_t)BK, (cuuint32_t)n_tile, 1};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(ptr), gdim, gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { set_error("conv_tc: cuTensorMapEncodeTiled(weight Kp=%d R=%d taps=%d n_tile=%d) failed (%d)", Kp, R, taps, n_tile, (int)r); return GDN_ECUDA; }
  return GDN_OK;
}

// the packed weights as [n_tile rows][8 columns] boxes without swizzle (COL mode: K groups of the resident weight operand)
static int make_weight_map8(CUtensorMap* m, const void* ptr, int Kp, int R, int taps, int n_tile) {
  EncodeTiledFn enc = get_encode();
  if (!enc) { set_error("conv_tc: cuTensorMapEncodeTiled unavailable"); return GDN_ECUDA; }
  cuuint64_t gdim[3] = {(cuuint64_t)Kp, (cuuint64_t)R, (cuuint64_t)taps};
  cuuint64_t gstride[2] = {(cuuint64_t)Kp * 2, (cuuint64_t)R * Kp * 2};
  cuuint32_t box[3] = {8, (cuuint32_t)n_tile, 1};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(ptr), gdim, gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { set_error("conv_tc: cuTensorMapEncodeTiled(weight groups Kp=%d R=%d taps=%d n_tile=%d) failed (%d)", Kp, R, taps, n_tile, (int)r); return GDN_ECUDA; }
  return GDN_OK;
}

// pixel tile shape: Wt = smallest power of two >= min(W, 128) (>= 8), Ht = 128 / Wt
static void tile_shape(int W, int* Wt, int* Ht) {
  int wt = 8;
  while (wt < W && wt < 128) wt <<= 1;
  *Wt = wt; *Ht = BM / wt;
}
static int pow2_cols(int n) { int c = 32; while (c < n) c <<= 1; return c; }
}  // namespace convtc
}  // namespace gdn

using namespace gdn;
using namespace gdn::convtc;

static bool g_halo_enabled = true;
static bool g_col_enabled = true;
/* test hook: the COL mode of the forward kernel (data gradient with a narrow gradient operand: DenseNet growth convolutions); returns the previous setting */
extern "C" int gdn_conv_tc_set_col(int enabled) { const int old = g_col_enabled; g_col_enabled = enabled != 0; return old; }
/* test hook: switch the halo-reuse variant of the forward kernel off (returns the previous setting) */
extern "C" int gdn_conv_tc_set_halo(int enabled) { const int old = g_halo_enabled; g_halo_enabled = enabled != 0; return old; }

extern "C" int gdn_conv_tc_init(void) {
  GDN_CHECK_CUDA(cudaFuncSetAttribute(conv_tc_fwd_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_LIMIT));
  GDN_CHECK_CUDA(cudaFuncSetAttribute(conv_tc_fwd_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_LIMIT));
  GDN_CHECK_CUDA(cudaFuncSetAttribute(conv_tc_fwd_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_LIMIT));
  GDN_CHECK_CUDA(cudaFuncSetAttribute(conv_tc_wgrad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_LIMIT));
  GDN_CHECK_CUDA(cudaFuncSetAttribute(conv_tc_wgrad_col_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_LIMIT));
  GDN_CHECK_CUDA(cudaFuncSetAttribute(conv_tc_wgrad_col_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_LIMIT));
  return GDN_OK;
}

extern "C" int gdn_pack_act_bf16(const float* x, int x_pitch, int x_c0, long long M, int C, uint16_t* hi, uint16_t* lo, const float* scale, const float* shift,
                                 int act, float slope, gdn_stream_t s) {
  GDN_CHECK_ARG(x && hi && M > 0 && C > 0 && x_pitch >= x_c0 + C);
  GDN_CHECK_ARG(((uintptr_t)hi & 15) == 0 && ((uintptr_t)lo & 15) == 0 && (scale == nullptr) == (shift == nullptr));
  const int Cp = (C + 7) & ~7;
  const long long total = M * (Cp / 8);
  const int blocks = (int)(cdiv(total, 256) < 16 * kNumSMs ? cdiv(total, 256) : 16 * kNumSMs);
  pack_act_kernel<<<blocks, 256, 0, as_stream(s)>>>(x + x_c0, x_pitch, M, C, Cp, reinterpret_cast<__nv_bfloat16*>(hi), reinterpret_cast<__nv_bfloat16*>(lo), scale, shift, act, slope, nullptr, 0, nullptr);
  GDN_CHECK_LAUNCH();
  return GDN_OK;
}

extern "C" int gdn_pack_actgrad_bf16(const float* dy, int dy_pitch, const float* y, int y_pitch, long long M, int C, uint16_t* hi, uint16_t* lo, int act, float slope,
                                     gdn_stream_t s) {
  GDN_CHECK_ARG(dy && y && hi && M > 0 && C > 0 && dy_pitch >= C && y_pitch >= C);
  GDN_CHECK_ARG(((uintptr_t)hi & 15) == 0 && ((uintptr_t)lo & 15) == 0);
  const int Cp = (C + 7) & ~7;
  const long long total = M * (Cp / 8);
  const int blocks = (int)(cdiv(total, 256) < 16 * kNumSMs ? cdiv(total, 256) : 16 * kNumSMs);
  pack_act_kernel<<<blocks, 256, 0, as_stream(s)>>>(dy, dy_pitch, M, C, Cp, reinterpret_cast<__nv_bfloat16*>(hi), reinterpret_cast<__nv_bfloat16*>(lo), nullptr, nullptr, act, slope,
                                                    y, y_pitch, nullptr);
  GDN_CHECK_LAUNCH();
  return GDN_OK;
}

extern "C" int gdn_pack_actgrad_bf16g(const float* dy, int dy_pitch, const uint16_t* y16, int y16_pitch, long long M, int C, uint16_t* hi, uint16_t* lo, int act,
                                      float slope, gdn_stream_t s) {
  GDN_CHECK_ARG(dy && y16 && hi && M > 0 && C > 0 && dy_pitch >= C && y16_pitch >= ((C + 7) & ~7) && y16_pitch % 8 == 0);
  GDN_CHECK_ARG(((uintptr_t)hi & 15) == 0 && ((uintptr_t)lo & 15) == 0 && ((uintptr_t)y16 & 15) == 0);
  const int Cp = (C + 7) & ~7;
  const long long total = M * (Cp / 8);
  const int blocks = (int)(cdiv(total, 256) < 16 * kNumSMs ? cdiv(total, 256) : 16 * kNumSMs);
  pack_act_kernel<<<blocks, 256, 0, as_stream(s)>>>(dy, dy_pitch, M, C, Cp, reinterpret_cast<__nv_bfloat16*>(hi), reinterpret_cast<__nv_bfloat16*>(lo), nullptr, nullptr, act, slope,
                                                    nullptr, y16_pitch, reinterpret_cast<const __nv_bfloat16*>(y16));
  GDN_CHECK_LAUNCH();
  return GDN_OK;
}

extern "C" size_t gdn_pack_weight_bf16_elems(int O, int I, int kh, int kw, int transposed) {
  const int R = transposed ? I : O, K = transposed ? O : I;
  return (size_t)kh * kw * R * ((K + 7) & ~7);
}
extern "C" int gdn_pack_weight_bf16(const float* w, int O, int I_total, int i_c0, int I, int kh, int kw, int transposed, uint16_t* hi, uint16_t* lo, gdn_stream_t s) {
  GDN_CHECK_ARG(w && hi && O > 0 && I > 0 && kh > 0 && kw > 0 && I_total >= i_c0 + I);
  const long long total = (long long)gdn_pack_weight_bf16_elems(O, I, kh, kw, transposed);
  const int blocks = (int)(cdiv(total, 256) < 8 * kNumSMs ? cdiv(total, 256) : 8 * kNumSMs);
  pack_weight_kernel<<<blocks, 256, 0, as_stream(s)>>>(w, O, I_total, i_c0, I, kh * kw, transposed, reinterpret_cast<__nv_bfloat16*>(hi), reinterpret_cast<__nv_bfloat16*>(lo));
  GDN_CHECK_LAUNCH();
  return GDN_OK;
}

extern "C" int gdn_conv2d_tc(const gdn_conv_tc_args* a, gdn_stream_t s) {
  GDN_CHECK_ARG(a && a->x_hi && a->w_hi && (a->y || a->y16));
  GDN_CHECK_ARG(a->B > 0 && a->Cin > 0 && a->Cout > 0 && a->kh > 0 && a->kw > 0 && a->kh * a->kw <= 9);
  GDN_CHECK_ARG(a->Hi > 0 && a->Wi > 0 && a->Ho > 0 && a->Wo > 0 && (a->stride == 1 || a->stride == 2));
  GDN_CHECK_ARG((!a->y || a->y_pitch >= a->y_c0 + a->Cout) && (!a->res || a->res_pitch >= a->res_c0 + a->Cout));
  GDN_CHECK_ARG(a->precision == GDN_PREC_BF16 || (a->precision == GDN_PREC_BF16X3 && a->x_lo && a->w_lo));
  GDN_CHECK_ARG(((uintptr_t)a->x_hi & 15) == 0 && ((uintptr_t)a->w_hi & 15) == 0 && ((uintptr_t)a->x_lo & 15) == 0 && ((uintptr_t)a->w_lo & 15) == 0);
  const int nsplit = a->precision == GDN_PREC_BF16X3 ? 3 : 1;
  const int Cp = (a->Cin + 7) & ~7;         // channel pitch of the packed activation AND K pitch of the packed weight
  const int taps = a->kh * a->kw;
  FwdParams p;
  p.y = a->y ? a->y + a->y_c0 : nullptr; p.y_pitch = a->y_pitch; p.bias = a->bias;
  p.y16 = reinterpret_cast<__nv_bfloat16*>(a->y16); p.y16_pitch = a->y16_pitch;
  p.res = a->res ? a->res + a->res_c0 : nullptr; p.res_pitch = a->res_pitch;
  p.alpha_ptr = a->alpha_ptr;
  const int groups = a->groups > 1 ? a->groups : 1;
  GDN_CHECK_ARG(a->B % groups == 0);
  p.imgs_per_group = a->B / groups; p.taps_total = a->kh * a->kw;
  p.B = a->B; p.Ho = a->Ho; p.Wo = a->Wo; p.Cout = a->Cout;
  p.kchunks = (int)cdiv(a->Cin, BK);
  {  // balanced N tiles: 368 output channels -> 2 x 192 instead of 256 + 112
    const int nt = (int)cdiv(a->Cout, 256);
    p.n_tile = (int)cdiv(cdiv(a->Cout, nt), 16) * 16;
  }
  p.nsplit = nsplit;
  p.tmem_cols = 2 * pow2_cols(p.n_tile);        // two accumulator buffers
  p.act = a->act; p.slope = a->slope;
  p.vec4 = (a->Cout % 4 == 0 && (!a->y || (a->y_pitch % 4 == 0 && a->y_c0 % 4 == 0 && ((uintptr_t)a->y & 15) == 0)) && (!a->bias || ((uintptr_t)a->bias & 15) == 0) &&
            (!a->res || (a->res_pitch % 4 == 0 && a->res_c0 % 4 == 0 && ((uintptr_t)a->res & 15) == 0))) ? 1 : 0;
  if (a->y16) GDN_CHECK_ARG(p.vec4 && a->y16_pitch % 4 == 0 && a->y16_pitch >= a->Cout && ((uintptr_t)a->y16 & 7) == 0);
  const int ring_budget = SMEM_LIMIT - 2048 - EPI_WARPS * EPI_STAGE_BYTES;     // bytes for the operand ring(s)
  const int sub_bytes = A_BYTES + p.n_tile * 128;
  // K-iterations per pipeline stage: narrow tiles (2*n_tile MMA cycles per K-iteration) amortise the mbarrier round trip over a group
  int kgroup = p.n_tile <= 64 ? (int)cdiv(256, p.n_tile) : 1;
  if (kgroup * sub_bytes > ring_budget / 3) kgroup = ring_budget / 3 / sub_bytes;      // at least three stages stay in flight
  if (kgroup < 1) kgroup = 1;
  p.kgroup = kgroup;
  const int stage_bytes = kgroup * sub_bytes;
  int stages = ring_budget / stage_bytes;
  if (stages > MAX_STAGES) stages = MAX_STAGES;
  p.stages = stages;
  p.halo_bo = 0;
  p.w_resident = 0;
  const size_t smem = (size_t)stages * stage_bytes + EPI_WARPS * EPI_STAGE_BYTES + 8 * (2 * MAX_STAGES + 10) + 1024;
  const int n_tiles = (int)cdiv(a->Cout, p.n_tile);
  cudaStream_t st = as_stream(s);
  int rc;

  // the (transposed ? Cout : Cin)-channel operand lives on the INPUT grid (Hi, Wi) of this launch
  const int nclass = (a->transposed && a->stride == 2) ? 4 : 1;
  for (int cls = 0; cls < nclass; ++cls) {
    const int ph = cls >> 1, pw = cls & 1;
    p.taps.n = 0;
    if (!a->transposed) {
      p.cs = a->stride; p.os = 1; p.oh0 = p.ow0 = 0; p.Hc = a->Ho; p.Wc = a->Wo;
      for (int kh = 0; kh < a->kh; ++kh)
        for (int kw = 0; kw < a->kw; ++kw) { int n = p.taps.n++; p.taps.dh[n] = kh - a->pad; p.taps.dw[n] = kw - a->pad; p.taps.widx[n] = kh * a->kw + kw; }
    } else if (a->stride == 1) {
      p.cs = 1; p.os = 1; p.oh0 = p.ow0 = 0; p.Hc = a->Ho; p.Wc = a->Wo;
      for (int kh = 0; kh < a->kh; ++kh)
        for (int kw = 0; kw < a->kw; ++kw) { int n = p.taps.n++; p.taps.dh[n] = a->pad - kh; p.taps.dw[n] = a->pad - kw; p.taps.widx[n] = kh * a->kw + kw; }
    } else {
      // output pixel (2i+ph, 2j+pw) <- dy pixel (i + (ph+pad-kh)/2, j + (pw+pad-kw)/2) for the taps where the division is exact
      p.cs = 1; p.os = 2; p.oh0 = ph; p.ow0 = pw;
      p.Hc = (a->Ho - ph + 1) / 2; p.Wc = (a->Wo - pw + 1) / 2;
      for (int kh = 0; kh < a->kh; ++kh)
        for (int kw = 0; kw < a->kw; ++kw) {
          const int eh = ph + a->pad - kh, ew = pw + a->pad - kw;
          if ((eh & 1) || (ew & 1)) continue;
          int n = p.taps.n++; p.taps.dh[n] = eh / 2; p.taps.dw[n] = ew / 2; p.taps.widx[n] = kh * a->kw + kw;
        }
      if (p.Hc <= 0 || p.Wc <= 0) continue;
      if (p.taps.n == 0) { set_error("gdn_conv2d_tc: parity class without taps is not supported (kernel smaller than stride)"); return GDN_EINVAL; }
    }
    tile_shape(p.Wc, &p.Wt, &p.Ht);
    p.wt_shift = 0;
    while ((1 << p.wt_shift) < p.Wt) ++p.wt_shift;
    p.tiles_w = (int)cdiv(p.Wc, p.Wt); p.tiles_h = (int)cdiv(p.Hc, p.Ht);
    p.n_tiles = n_tiles;
    const long long total = (long long)a->B * p.tiles_h * p.tiles_w * n_tiles;
    GDN_CHECK_ARG(total < (1ll << 31));
    p.total_tiles = (int)total;
    // halo mode: stride-1 3x3 neighbourhood, one-row tiles of 128 pixels, narrow output tile (otherwise the MMA, not L2, is the limit)
    // measured on B200 (tools/bench_conv.py): the halo box also pays off for wider output tiles when the K extent is short -- one channel
    // chunk (DenseBlock data gradients 24 -> C, n_tile <= 160: 0.15 -> 0.11 ms) or three chunks with 128 < n_tile <= 192 (DANet fuse
    // data gradients 184 -> 368: 0.46 -> 0.38 ms); it loses for 128 -> 128 and 368 -> 184, which stay on the per-tap variant
    const bool halo_shape = p.n_tile <= 64 || (p.kchunks == 1 && p.n_tile <= 160) || (p.kchunks <= 3 && p.n_tile > 128 && p.n_tile <= 192);
    bool halo = g_halo_enabled && p.cs == 1 && p.os == 1 && p.Wt == BM && p.Ht == 1 && p.taps.n == 9 && halo_shape;
    for (int tp = 0; tp < p.taps.n && halo; ++tp) halo = p.taps.dh[tp] >= -1 && p.taps.dh[tp] <= 1 && p.taps.dw[tp] >= -1 && p.taps.dw[tp] <= 1;
    CUtensorMap mxh, mxl, mwh, mwl;
    // COL mode: narrow-operand data gradient (see the kernel): one-row tiles of 128 pixels, one output-channel tile, bf16
    const size_t smem_col = (size_t)2 * WR_COL_BYTES + (size_t)3 * COL_KGROUPS * p.n_tile * 16 + EPI_WARPS * EPI_STAGE_BYTES + 8 * (2 * MAX_STAGES + 10) + 1024;
    const bool col = g_col_enabled && a->transposed && a->stride == 1 && a->kh == 3 && a->kw == 3 && a->pad == 1 && nsplit == 1 && groups == 1 && a->Cin <= 24 &&
                     p.Wt == BM && p.Ht == 1 && n_tiles == 1 && smem_col <= (size_t)SMEM_LIMIT;
    if (col) {
      FwdParams pc = p;
      pc.col_k = a->Cin; pc.stages = 1;
      if ((rc = make_act_map8(&mxh, a->x_hi, Cp, a->Wi, a->Hi, a->B, BM + 2, 1)) != GDN_OK) return rc;
      if ((rc = make_weight_map8(&mwh, a->w_hi, Cp, a->Cout, taps * groups, p.n_tile)) != GDN_OK) return rc;
      const int gridc = p.total_tiles < kNumSMs ? p.total_tiles : kNumSMs;
      conv_tc_fwd_kernel<2><<<gridc, FWD_THREADS, smem_col, st>>>(mxh, mxh, mwh, mwh, pc);
      GDN_CHECK_LAUNCH();
      continue;
    }
    const int bw = halo ? HALO_W : p.Wt, bh = halo ? HALO_ROWS : p.Ht;
    if ((rc = make_act_map(&mxh, a->x_hi, Cp, a->Wi, a->Hi, a->B, bw, bh, p.cs)) != GDN_OK) return rc;
    mxl = mxh;
    if (nsplit == 3 && (rc = make_act_map(&mxl, a->x_lo, Cp, a->Wi, a->Hi, a->B, bw, bh, p.cs)) != GDN_OK) return rc;
    if ((rc = make_weight_map(&mwh, a->w_hi, Cp, a->Cout, taps * groups, p.n_tile)) != GDN_OK) return rc;
    mwl = mwh;
    if (nsplit == 3 && (rc = make_weight_map(&mwl, a->w_lo, Cp, a->Cout, taps * groups, p.n_tile)) != GDN_OK) return rc;
    const int grid = p.total_tiles < kNumSMs ? p.total_tiles : kNumSMs;     // persistent: one CTA per SM
    if (halo) {
      FwdParams ph = p;
      const int b_bytes = p.n_tile * 128, w_budget = ring_budget - HALO_SLOTS * HALO_SLOT;
      const long long w_all = (long long)p.kchunks * nsplit * 9 * b_bytes;
      ph.w_resident = (n_tiles == 1 && groups == 1 && w_all <= w_budget && w_all < (1 << 20) && p.kchunks * nsplit <= MAX_STAGES) ? 1 : 0;
      ph.kgroup = 9 * b_bytes * 2 <= w_budget ? 9 : (3 * b_bytes * 2 <= w_budget ? 3 : 1);   // taps per weight stage: all 9, one filter row, or one
      ph.stages = w_budget / (ph.kgroup * b_bytes);
      if (ph.stages > MAX_STAGES) ph.stages = MAX_STAGES;
      { const char* e = getenv("GDN_HALO_BO"); ph.halo_bo = e ? atoi(e) : 0; }   // measured on B200: the swizzle phase comes from the absolute smem address, base_offset must stay 0
      if (ph.w_resident) { ph.kgroup = 9; ph.stages = p.kchunks * nsplit; if (ph.stages > MAX_STAGES) ph.stages = MAX_STAGES; }   // ring area = the resident weight block
      const size_t smem_h = (size_t)HALO_SLOTS * HALO_SLOT + (ph.w_resident ? (size_t)w_all : (size_t)ph.stages * ph.kgroup * b_bytes) + EPI_WARPS * EPI_STAGE_BYTES +
                            8 * (2 * MAX_STAGES + 10) + 1024;
      conv_tc_fwd_kernel<1><<<grid, FWD_THREADS, smem_h, st>>>(mxh, mxl, mwh, mwl, ph);
    } else {
      conv_tc_fwd_kernel<0><<<grid, FWD_THREADS, smem, st>>>(mxh, mxl, mwh, mwl, p);
    }
    GDN_CHECK_LAUNCH();
  }
  return GDN_OK;
}

static void wgrad_plan(const gdn_wgrad_tc_args* a, WgParams* p) {
  const int taps = a->kh * a->kw;
  p->direct = nullptr; p->direct_pitch = 0; p->Cin = a->Cin; p->scale_ptr = nullptr;
  p->Cout = a->Cout; p->taps_total = taps; p->kw = a->kw; p->pad = a->pad; p->cs = a->stride;
  p->ci_tile = a->Cin > 64 ? 128 : 64;
  p->ci_tiles = (int)cdiv(a->Cin, p->ci_tile);
  p->cin_w = p->ci_tiles * p->ci_tile;
  p->co_tiles = (int)cdiv(a->Cout, 128);
  p->taps_per_cta = a->kw;                       // one filter row per CTA: kw accumulators of ci_tile columns
  p->tap_groups = a->kh;
  p->tmem_cols = pow2_cols(p->taps_per_cta * p->ci_tile);
  tile_shape(a->Wo, &p->Wt, &p->Ht);
  p->tiles_w = (int)cdiv(a->Wo, p->Wt); p->tiles_h = (int)cdiv(a->Ho, p->Ht);
  p->tiles_total = a->B * p->tiles_h * p->tiles_w;
  const int units = p->co_tiles * p->ci_tiles * p->tap_groups;
  // one CTA per SM (128-198 KB of shared memory): the grid is a whole number of waves -- 2 * 148 / units rounded DOWN (rounded up, 300 CTAs on 148 SMs
  // ran as three waves: ncu round 2, the dense-block growth convolutions)
  int splits = (2 * kNumSMs) / units;
  if (splits < 1) splits = 1;
  p->x_slots = p->ci_tile == 64 ? 8 : 4;         // 64-channel tiles: 16 KB slots, eight of them (the 128-channel tile's four 32 KB slots fill the budget)
  const int max_splits = p->tiles_total / 8 > 0 ? p->tiles_total / 8 : 1;
  if (splits > max_splits) splits = max_splits;
  p->tiles_per_split = (int)cdiv(p->tiles_total, splits);
  if (a->groups > 1) p->tiles_per_split = p->tiles_total / a->groups;     // grouped: exactly one split per sample
  p->nsplit = a->precision == GDN_PREC_BF16X3 ? 3 : 1;
}
static int wgrad_splits(const WgParams* p) { return (int)cdiv(p->tiles_total, p->tiles_per_split); }

// Narrow outputs (the DenseNet growth convolutions, generator.py:34: Cout = 24): with the output channels as the UMMA M dimension 24 of 128 accumulator
// lanes do work.  dW[co][tap][ci] = sum_q dy[q][co] x[q + tap][ci] = sum_p x[p][ci] dy[p - tap][co] is the SAME kernel with the operands' roles swapped
// (x as the "gradient" operand: M = input channels, dy as the shifted operand: N = 64 >= Cout) and the filter taps mirrored, for stride-1 "same"
// convolutions (equal grids: the zero padding of one operand's shift is the other's).  The reduction pass transposes the partial sums back.
static int g_wgrad_swap = 0;      // measured (tools/profile_wgrad.py): no faster than the direct launch (0.187 vs 0.189 ms at C_in 136, slower at C_in 64) --
                                  // these weight gradients are bound by the operand ring's latency, not by the MMA; kept behind the hook, tested
extern "C" int gdn_conv_tc_set_wgrad_swap(int enabled) { const int old = g_wgrad_swap; g_wgrad_swap = enabled ? 1 : 0; return old; }
static bool wgrad_swap_ok(const gdn_wgrad_tc_args* a) {
  return g_wgrad_swap && a->groups <= 1 && a->stride == 1 && a->Cout <= 32 && a->Cin >= 64 && a->Ho == a->Hi && a->Wo == a->Wi && a->kh == a->kw &&
         2 * a->pad == a->kh - 1 && (a->kh & 1);
}
static gdn_wgrad_tc_args wgrad_swapped(const gdn_wgrad_tc_args* a) {
  gdn_wgrad_tc_args b = *a;
  b.dy_hi = a->x_hi; b.dy_lo = a->x_lo; b.x_hi = a->dy_hi; b.x_lo = a->dy_lo;
  b.Cout = a->Cin; b.Cin = a->Cout;
  return b;
}

// conv_tc_wgrad_col_kernel applies to 3x3 stride-1 "same" convolutions with at most 24 output channels (9 taps x 3 groups of 8 = 216 <= 256 accumulator rows)
static int g_wgrad_col = 1;
extern "C" int gdn_conv_tc_set_wgrad_col(int enabled) { const int old = g_wgrad_col; g_wgrad_col = enabled ? 1 : 0; return old; }
static bool wgrad_col_ok(const gdn_wgrad_tc_args* a) {
  return g_wgrad_col && a->groups <= 1 && a->precision == GDN_PREC_BF16 && a->kh == 3 && a->kw == 3 && a->stride == 1 && a->pad == 1 && a->Ho == a->Hi && a->Wo == a->Wi &&
         a->Cout <= 24 && a->Cin >= 16 && a->Cin <= 192;
}
static int g_wgrad_col_row = 1;    // test hook: the one-image-row variant (3 * G boxes of 130 pixels) where the tile is a 128-pixel row
extern "C" int gdn_conv_tc_set_wgrad_col_row(int enabled) { const int old = g_wgrad_col_row; g_wgrad_col_row = enabled ? 1 : 0; return old; }
static void wgrad_col_plan(const gdn_wgrad_tc_args* a, WcParams* p, int* ctas, bool* row) {
  p->Cout = a->Cout; p->G = (a->Cout + 7) / 8; p->ngroups = 9 * p->G;
  p->nbox = (int)cdiv(a->Cin, 64); p->N = (int)cdiv(a->Cin, 16) * 16;
  tile_shape(a->Wo, &p->Wt, &p->Ht);
  *row = g_wgrad_col_row && p->Wt == BM && p->Ht == 1 && 3 * p->N + 32 <= 512;
  p->tmem_cols = pow2_cols((*row ? 3 : 2) * p->N + 32);
  p->tiles_w = (int)cdiv(a->Wo, p->Wt); p->tiles_h = (int)cdiv(a->Ho, p->Ht);
  p->tiles_total = a->B * p->tiles_h * p->tiles_w;
  *ctas = p->tiles_total < kNumSMs ? p->tiles_total : kNumSMs;
  p->lbo = 128; p->sbo = *row ? WR_SLOT_BYTES : WC_GROUP_BYTES;
}
static int wgrad_col_launch(const gdn_wgrad_tc_args* a, cudaStream_t st) {
  WcParams p; int ctas; bool row;
  wgrad_col_plan(a, &p, &ctas, &row);
  GDN_CHECK_ARG(((uintptr_t)a->ws & 15) == 0 && p.tmem_cols <= 512);
  p.ws = a->ws;
  const int Cop = (a->Cout + 7) & ~7, Cip = (a->Cin + 7) & ~7;
  CUtensorMap mdy, mx;
  int rc;
  if ((rc = make_act_map8(&mdy, a->dy_hi, Cop, a->Wo, a->Ho, a->B, row ? BM + 2 : p.Wt, p.Ht)) != GDN_OK) return rc;
  if ((rc = make_act_map(&mx, a->x_hi, Cip, a->Wi, a->Hi, a->B, p.Wt, p.Ht, 1)) != GDN_OK) return rc;
  const size_t smem = (size_t)WC_STAGES * ((row ? WR_COL_BYTES : WC_COL_BYTES) + p.nbox * A_BYTES) + 8 * (2 * WC_STAGES + 2) + 16 + 1024;
  GDN_CHECK_ARG(smem <= (size_t)SMEM_LIMIT);
  if (row) conv_tc_wgrad_col_kernel<true><<<ctas, NTHREADS, smem, st>>>(mdy, mx, p);
  else conv_tc_wgrad_col_kernel<false><<<ctas, NTHREADS, smem, st>>>(mdy, mx, p);
  GDN_CHECK_LAUNCH();
  const long long total = (long long)a->Cout * 9 * a->Cin;
  const int blocks = (int)(cdiv(total, 256) < 4 * kNumSMs ? cdiv(total, 256) : 4 * kNumSMs);
  wgrad_col_reduce_kernel<<<blocks, 256, 0, st>>>(a->ws, ctas, a->Cout, a->Cin, p.N, a->out, a->out_cin_total, a->out_c0, a->accumulate, a->scale, a->scale_ptr);
  GDN_CHECK_LAUNCH();
  return GDN_OK;
}

extern "C" size_t gdn_conv2d_wgrad_tc_ws_bytes(const gdn_wgrad_tc_args* a) {
  if (a->groups > 1) return 0;
  if (wgrad_col_ok(a)) { WcParams pc; int ctas; bool row; wgrad_col_plan(a, &pc, &ctas, &row); return (size_t)ctas * 9 * a->Cout * pc.N * sizeof(float); }
  WgParams p;
  if (wgrad_swap_ok(a)) {
    const gdn_wgrad_tc_args b = wgrad_swapped(a);
    wgrad_plan(&b, &p);
    return (size_t)wgrad_splits(&p) * b.Cout * p.taps_total * p.cin_w * sizeof(float);
  }
  wgrad_plan(a, &p);
  return (size_t)wgrad_splits(&p) * a->Cout * p.taps_total * p.cin_w * sizeof(float);
}

extern "C" int gdn_conv2d_wgrad_tc(const gdn_wgrad_tc_args* orig, gdn_stream_t s) {
  GDN_CHECK_ARG(orig && orig->dy_hi && orig->x_hi && orig->out && (orig->ws || orig->groups > 1));
  GDN_CHECK_ARG(orig->out_cin_total >= orig->out_c0 + orig->Cin);
  if (wgrad_col_ok(orig)) {
    if (orig->ws_bytes < gdn_conv2d_wgrad_tc_ws_bytes(orig)) { set_error("gdn_conv2d_wgrad_tc: workspace too small"); return GDN_EWORKSPACE; }
    return wgrad_col_launch(orig, as_stream(s));
  }
  const bool swapped = wgrad_swap_ok(orig);
  const gdn_wgrad_tc_args sw = swapped ? wgrad_swapped(orig) : *orig;
  const gdn_wgrad_tc_args* a = &sw;
  GDN_CHECK_ARG(a->groups <= 1 || (a->groups == a->B && a->kh == 1 && a->kw == 1 && a->Cin % 4 == 0 && !a->accumulate && ((uintptr_t)a->out & 15) == 0));
  GDN_CHECK_ARG(a->B > 0 && a->Cin > 0 && a->Cout > 0 && a->kh > 0 && a->kw > 0 && a->kw <= 3 && a->kh <= 3 && (a->stride == 1 || a->stride == 2));
  GDN_CHECK_ARG(a->precision == GDN_PREC_BF16 || (a->precision == GDN_PREC_BF16X3 && a->dy_lo && a->x_lo));
  WgParams p;
  wgrad_plan(a, &p);
  const int splits = wgrad_splits(&p);
  if (a->ws_bytes < gdn_conv2d_wgrad_tc_ws_bytes(orig)) { set_error("gdn_conv2d_wgrad_tc: workspace too small"); return GDN_EWORKSPACE; }
  GDN_CHECK_ARG(((uintptr_t)a->ws & 15) == 0);
  p.ws = a->ws;
  if (a->groups > 1) { p.direct = a->out; p.direct_pitch = a->Cin; p.scale_ptr = a->scale_ptr; }
  const int Cop = (a->Cout + 7) & ~7, Cip = (a->Cin + 7) & ~7;
  CUtensorMap mdh, mdl, mxh, mxl;
  int rc;
  if ((rc = make_act_map(&mdh, a->dy_hi, Cop, a->Wo, a->Ho, a->B, p.Wt, p.Ht, 1)) != GDN_OK) return rc;
  mdl = mdh;
  if (p.nsplit == 3 && (rc = make_act_map(&mdl, a->dy_lo, Cop, a->Wo, a->Ho, a->B, p.Wt, p.Ht, 1)) != GDN_OK) return rc;
  if ((rc = make_act_map(&mxh, a->x_hi, Cip, a->Wi, a->Hi, a->B, p.Wt, p.Ht, p.cs)) != GDN_OK) return rc;
  mxl = mxh;
  if (p.nsplit == 3 && (rc = make_act_map(&mxl, a->x_lo, Cip, a->Wi, a->Hi, a->B, p.Wt, p.Ht, p.cs)) != GDN_OK) return rc;
  const size_t smem = (size_t)WG_DY_SLOTS * WG_DY_BYTES + (size_t)p.x_slots * (p.ci_tile / 64) * A_BYTES + 8 * (2 * WG_DY_SLOTS + 2 * WG_X_SLOTS + 2) + 1024;
  dim3 grid((unsigned)(p.co_tiles * p.ci_tiles * p.tap_groups), (unsigned)splits);
  GDN_CHECK_ARG(grid.y <= 65535);
  cudaStream_t st = as_stream(s);
  conv_tc_wgrad_kernel<<<grid, NTHREADS, smem, st>>>(mdh, mdl, mxh, mxl, p);
  GDN_CHECK_LAUNCH();
  if (a->groups > 1) return GDN_OK;
  const long long total = (long long)a->Cout * p.taps_total * a->Cin;
  const int blocks = (int)(cdiv(total, 256) < 4 * kNumSMs ? cdiv(total, 256) : 4 * kNumSMs);
  wgrad_tc_reduce_kernel<<<blocks, 256, 0, st>>>(a->ws, splits, orig->Cout, orig->Cin, p.taps_total, p.cin_w, a->out, a->out_cin_total, a->out_c0, a->accumulate, a->scale,
                                                 a->scale_ptr, swapped ? 1 : 0);
  GDN_CHECK_LAUNCH();
  return GDN_OK;
}
