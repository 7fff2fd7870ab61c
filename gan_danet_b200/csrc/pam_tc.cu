// Position-attention core as ONE fused flash-style kernel for sm_100a:
//   TMA (cp.async.bulk.tensor) -> shared memory -> tcgen05.mma (kind::f16, fp32 accumulate in TMEM) -> online softmax
//   on tcgen05.ld'ed score tiles -> P (bf16, written back to TMEM) -> tcgen05.mma P.V (A operand from TMEM) -> epilogue
//   y = gamma * O / l + x,  lse.
// The N x N attention map of the reference (torch.bmm + softmax + torch.bmm at /root/reference/models/generator.py:115-122)
// never leaves the SM.
//
// One CTA = one sample x one tile of 128 query positions; it walks all N/64 key tiles of 64 keys.
//   warp 0      : TMA producer for Q and the K ring (6 stages of 64 keys)
//   warp 3      : TMA producer for the V ring (6 stages of 64 keys x 192 channels, row-major: V is the MN-major B operand)
//   warp 1      : tcgen05.mma issuer for  O += P_t V_t        (warp-uniform loop, one elected lane issues)
//   warp 2      : TMEM allocator, then tcgen05.mma issuer for  S[t&3] = Q K_t^T
//   warps 4-7   : softmax group 0 -- EVEN key tiles; thread = query row = TMEM lane, all 64 key columns of the tile in registers
//   warps 8-11  : softmax group 1 -- ODD key tiles
// The two groups ping-pong: while one is in its MUFU-bound exp phase the other does its latency-bound phase (row maximum, P
// store, fences), so the exp units and the tensor pipe stay busy.  A row's maximum is thread-local (no cross-group exchange);
// what the groups share is the per-row REFERENCE maximum in shared memory, handed over warp-to-warp (mbarrier) in tile order.
// The reference starts 2^32 above the first tile's maximum and is raised lazily (only when a tile exceeds it by 2^64: P is bf16, so
// its 8 exponent bits cover the whole window); the raising thread rescales its row of O in TMEM after the previous P.V has landed.
// With a reference that tracked the maximum closely (fp16 P, window 2^15) the real activations of the network (|logit| ~ 100)
// raised it in 8-14 % of the tiles and each raise stalls on the tensor pipe (ncu, profiles/r01_pam_fwd_notes.txt).
// V carries a channel of ones at column C: the softmax denominator l is the P.V product's column C, i.e. the sum of the SAME rounded
// bf16 weights the numerator uses (a one-hot row returns its value row exactly), and no row sum is accumulated in the softmax threads.
// TMEM columns: four score buffers S[u] = [64u, 64u+64), O [256,448).  P never touches shared memory: a group writes its fp16 P
// tile back into the first 32 columns of the tile's own score buffer (tcgen05.st) and the P.V product takes its A operand from
// TMEM (with P staged in shared memory the kernel is bound by shared-memory bandwidth: ncu showed tensor-pipe + LSU + TMA
// wavefronts ~ 1600 of 1950 cycles per 128 keys).  Each group owns two score buffers, so Q K_{t+2}^T of its NEXT tile is
// already in TMEM when it finishes tile t (issued right after P_{t-2} V_{t-2}): the tensor-pipe latency is off the softmax
// groups' critical path.  Q K_{t+4}^T goes into the buffer P_t was read from, after P_t V_t has completed (mbarrier).  The two
// products have separate issuing warps: a single warp's serial instruction stream (waits, descriptor adds, issue, commits) was
// the kernel's critical path at ~900 cycles per 64 keys.
// Epilogue: O/l is staged through shared memory (the drained V ring) so that x is read and o, y are written as whole rows.
// Operands: Q, K fp16 (bf16 logits are 8x worse), d zero-padded to 32; P, V bf16, C zero-padded to 192 (C < 192).
#include <cuda_bf16.h>
#include "tc_common.cuh"

namespace gdn {
namespace pamtc {
using namespace gdn::tc;

constexpr int TQ = 128, TK = 64, DPAD = 32, CPAD = 192;
constexpr int KS = 6, VS = 6;
constexpr int Q_CHUNK = TQ * DPAD * 2;          // 8 KB, 64-byte rows, SWIZZLE_64B
constexpr int K_CHUNK = TK * DPAD * 2;          // 4 KB
constexpr int V_CHUNK = TK * 128;               // 8 KB: [64 keys][64 channels] fp16, 128-byte rows, SWIZZLE_128B
constexpr int V_BYTES = (CPAD / 64) * V_CHUNK;  // 24 KB per key tile
constexpr int STG_STRIDE = CPAD + 4;            // floats per staged output row (conflict-free float4 column writes)
// Shared-memory plan.  SPLIT = 1 (precision GDN_PREC_FP16X3): the logit operands are fp16 hi + lo pairs (q = q_hi + q_lo exactly to 22 mantissa bits),
// stored as two 32-column chunks per tile; S = q_hi k_hi^T + q_lo k_hi^T + q_hi k_lo^T accumulates in the same TMEM buffer (three pairs of K = 16
// MMAs instead of one).  The reference has NO 1/sqrt(d) scale (generator.py:115-118): logits reach +-100, where a single fp16 product is off by
// 0.05 -- a 5 % error of the softmax weight; the split brings the logit error to ~1e-5.
template <int SPLIT>
struct FwdSmem {
  static constexpr int Q_BYTES = Q_CHUNK * (1 + SPLIT);
  static constexpr int K_BYTES = K_CHUNK * (1 + SPLIT);
  static constexpr int OFF_Q = 0;
  static constexpr int OFF_K = OFF_Q + Q_BYTES;
  static constexpr int OFF_V = OFF_K + KS * K_BYTES;     // 32768 / 65536 (1024-aligned)
  static constexpr int OFF_BAR = OFF_V + VS * V_BYTES;
  static constexpr int OFF_MREF = OFF_BAR + 512;         // float [128]: shared reference maximum per row (log2 units)
  static constexpr int OFF_LS = OFF_MREF + 128 * 4;      // float [2][128]
  static constexpr int OFF_TMEM = OFF_LS + 2 * 128 * 4;  // uint32 tmem base
  static constexpr int SMEM_BYTES = OFF_TMEM + 16 + 1024; // + alignment slack
  static_assert(OFF_V % 1024 == 0, "swizzle atoms");
  static_assert(TQ * STG_STRIDE * 4 <= VS * V_BYTES, "epilogue staging fits in the V ring");
  static_assert(SMEM_BYTES <= 227 * 1024, "shared memory budget");
};
constexpr int NTHREADS = 384;
constexpr uint32_t COL_O = 256;
constexpr float LOG2E = 1.4426950408889634f;
constexpr float LN2 = 0.6931471805599453f;
constexpr float RESCALE_TAU = 64.f;    // raise the reference when a tile maximum exceeds it by 2^64 (P is bf16: 8 exponent bits)
constexpr float REF_MARGIN = 32.f;     // ... and then put it 2^32 ABOVE that maximum: with real activations (|logit| ~ 100) a reference that
                                       // tracks the maximum closely is raised in ~10 % of the tiles, each time stalling on the tensor pipe

// barrier slots (8 bytes each) inside OFF_BAR
enum { BAR_Q = 0, BAR_KFULL = 1, BAR_KEMPTY = BAR_KFULL + KS, BAR_VFULL = BAR_KEMPTY + KS, BAR_VEMPTY = BAR_VFULL + VS,
       BAR_SFULL = BAR_VEMPTY + VS, BAR_PFULL = BAR_SFULL + 4, BAR_PVDONE = BAR_PFULL + 4, BAR_MREF = BAR_PVDONE + 4, BAR_COUNT = BAR_MREF + 8 };
static_assert(BAR_COUNT * 8 <= 512, "barrier area");

struct FwdParams {
  const float* x; int x_pitch; const float* gamma;
  float* o; float* y; int y_pitch; float* lse;
  int B, N, C, tiles_per_sample;
  __nv_bfloat16* y16; int y16_pitch;   // optional: y also (or, with y == NULL, only) as bf16 into a wider packed operand buffer (the DANet fuse convolution's input)
  int n_pv;      // MMA N of the P.V product: round_up(C + 1, 16) <= CPAD (C valid channels + the channel of ones); columns beyond it are never accumulated
};

__device__ __forceinline__ uint64_t smem_desc_mn(uint32_t addr, uint32_t lbo_bytes, uint32_t sbo_bytes, uint32_t layout_type) {
  return (uint64_t)((addr >> 4) & 0x3FFF) | ((uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16) | ((uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32) | ((uint64_t)1 << 46) |
         ((uint64_t)layout_type << 61);
}

// Y16 = 1: y is written only as a bf16 column block (FwdParams::y16).  A separate instantiation: with the bf16 store as a run-time branch of the
// one kernel, the fp32-store launches ran 4 % slower inside the training step (846 -> 813 TFLOP/s, measured three times each).
template <int Y16, int SPLIT>
__global__ void __launch_bounds__(NTHREADS, 1)
pam_flash_fwd_kernel(const __grid_constant__ CUtensorMap mapQ, const __grid_constant__ CUtensorMap mapK, const __grid_constant__ CUtensorMap mapV, const FwdParams p) {
  using SM = FwdSmem<SPLIT>;
  constexpr int Q_BYTES = SM::Q_BYTES, K_BYTES = SM::K_BYTES, OFF_Q = SM::OFF_Q, OFF_K = SM::OFF_K, OFF_V = SM::OFF_V, OFF_BAR = SM::OFF_BAR,
                OFF_MREF = SM::OFF_MREF, OFF_TMEM = SM::OFF_TMEM;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* sm = smem_raw + (base - smem_u32(smem_raw));
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int sample = blockIdx.x / p.tiles_per_sample, qtile = blockIdx.x % p.tiles_per_sample;
  const int T = p.N / TK;
  auto bar = [&](int i) { return base + OFF_BAR + 8 * i; };
  volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(sm + OFF_TMEM);

  if (threadIdx.x == 0) {
    mbar_init(bar(BAR_Q), 1);
    for (int i = 0; i < KS; ++i) { mbar_init(bar(BAR_KFULL + i), 1); mbar_init(bar(BAR_KEMPTY + i), 1); }
    for (int i = 0; i < VS; ++i) { mbar_init(bar(BAR_VFULL + i), 1); mbar_init(bar(BAR_VEMPTY + i), 1); }
    for (int i = 0; i < 4; ++i) { mbar_init(bar(BAR_SFULL + i), 1); mbar_init(bar(BAR_PFULL + i), 4); mbar_init(bar(BAR_PVDONE + i), 1); }
    for (int i = 0; i < 8; ++i) mbar_init(bar(BAR_MREF + i), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(base + OFF_TMEM), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;

  if (warp == 0) {
    if (lane == 0) {   // ---- Q + K producer
      mbar_expect_tx(bar(BAR_Q), Q_BYTES);
      tma_load_2d(base + OFF_Q, &mapQ, bar(BAR_Q), 0, sample * p.N + qtile * TQ);
      if (SPLIT) tma_load_2d(base + OFF_Q + Q_CHUNK, &mapQ, bar(BAR_Q), DPAD, sample * p.N + qtile * TQ);          // the lo halves: columns [32, 64)
      for (int t = 0; t < T; ++t) {
        const int st = t % KS;
        if (t >= KS) mbar_wait(bar(BAR_KEMPTY + st), ((t / KS) - 1) & 1);
        mbar_expect_tx(bar(BAR_KFULL + st), K_BYTES);
        tma_load_2d(base + OFF_K + st * K_BYTES, &mapK, bar(BAR_KFULL + st), 0, sample * p.N + t * TK);
        if (SPLIT) tma_load_2d(base + OFF_K + st * K_BYTES + K_CHUNK, &mapK, bar(BAR_KFULL + st), DPAD, sample * p.N + t * TK);
      }
    }
  } else if (warp == 3) {
    if (lane == 0) {   // ---- V producer: three 64-channel chunks per key tile
      for (int t = 0; t < T; ++t) {
        const int st = t % VS;
        if (t >= VS) mbar_wait(bar(BAR_VEMPTY + st), ((t / VS) - 1) & 1);
        mbar_expect_tx(bar(BAR_VFULL + st), V_BYTES);
#pragma unroll
        for (int c = 0; c < CPAD / 64; ++c)
          tma_load_2d(base + OFF_V + st * V_BYTES + c * V_CHUNK, &mapV, bar(BAR_VFULL + st), c * 64, sample * p.N + t * TK);
      }
    }
  } else if (warp == 1) {
    // ---- P.V issuer: the whole warp runs the loop (uniform control flow and descriptors), one elected lane issues
    // A = P and B = V are bf16 (P needs bf16's exponent range for the lazy reference; a bf16 x fp16 operand pair raises an illegal-instruction
    // fault on sm_100a: kind::f16 wants ONE format for A and B -- tried in round 2); bit 16: B is MN-major.  N = n_pv: C = 160 needs 176 of the 192 padded channel columns (measured: no change, 806 TFLOP/s either way -- the tensor pipe is not this kernel's limiter)
    const uint32_t IDESC_PV = idesc_f16(TQ, p.n_pv) | (1u << 7) | (1u << 10) | (1u << 16);
    const uint64_t vd_base = smem_desc_mn(base + OFF_V, V_CHUNK, 1024, LAYOUT_SW128);     // 64-channel groups 8 KB apart, 8-key groups 1 KB apart
    const uint32_t tmem_o = tmem + COL_O;
    int st = 0; uint32_t ph = 0;
    uint64_t vd = vd_base;
    for (int t = 0; t < T; ++t) {
      const int u = t & 3;
      mbar_wait_spin(bar(BAR_PFULL + u), (t >> 2) & 1);     // P_t is in TMEM (columns [0,32) of S[u])
      mbar_wait_spin(bar(BAR_VFULL + st), ph);
      tc_fence_after();
      const uint32_t pa = tmem + u * TK;
      if (elect_one()) {
        // A: 16 keys = 8 TMEM columns of packed fp16 pairs; B: 16 keys (2 KB) per step
        if (t == 0) umma_f16_ts_i<0>(tmem_o, pa, vd, IDESC_PV); else umma_f16_ts_i<1>(tmem_o, pa, vd, IDESC_PV);
        umma_f16_ts_i<1>(tmem_o, pa + 8, vd + 128, IDESC_PV);
        umma_f16_ts_i<1>(tmem_o, pa + 16, vd + 256, IDESC_PV);
        umma_f16_ts_i<1>(tmem_o, pa + 24, vd + 384, IDESC_PV);
        tc_commit(bar(BAR_VEMPTY + st));
        tc_commit(bar(BAR_PVDONE + u));
      }
      __syncwarp();
      vd += V_BYTES >> 4;
      if (++st == VS) { st = 0; ph ^= 1; vd = vd_base; }
    }
  } else if (warp == 2) {
    // ---- Q.K^T issuer (after its TMEM allocation duty): S[t&3] may be overwritten once P_{t-4} V_{t-4} has read P_{t-4} from it
    constexpr uint32_t IDESC_QK = idesc_f16(TQ, TK);
    const uint64_t qd0 = smem_desc(base + OFF_Q, 512, LAYOUT_SW64), qd1 = smem_desc(base + OFF_Q + 32, 512, LAYOUT_SW64);
    const uint64_t kd_base = smem_desc(base + OFF_K, 512, LAYOUT_SW64);
    int st = 0; uint32_t ph = 0;
    uint64_t kd = kd_base;
    mbar_wait_spin(bar(BAR_Q), 0);
    for (int t = 0; t < T; ++t) {
      const int u = t & 3;
      mbar_wait_spin(bar(BAR_KFULL + st), ph);
      if (t >= 4) mbar_wait_spin(bar(BAR_PVDONE + u), ((t >> 2) - 1) & 1);
      tc_fence_after();
      const uint32_t d = tmem + u * TK;
      if (elect_one()) {
        if (SPLIT) {      // the two small cross terms first, the leading term last
          umma_f16_i<0>(d, qd0 + (Q_CHUNK >> 4), kd, IDESC_QK);                        // q_lo k_hi^T
          umma_f16_i<1>(d, qd1 + (Q_CHUNK >> 4), kd + 2, IDESC_QK);
          umma_f16_i<1>(d, qd0, kd + (K_CHUNK >> 4), IDESC_QK);                        // q_hi k_lo^T
          umma_f16_i<1>(d, qd1, kd + (K_CHUNK >> 4) + 2, IDESC_QK);
          umma_f16_i<1>(d, qd0, kd, IDESC_QK);                                         // q_hi k_hi^T
        } else {
          umma_f16_i<0>(d, qd0, kd, IDESC_QK);
        }
        umma_f16_i<1>(d, qd1, kd + 2, IDESC_QK);
        tc_commit(bar(BAR_KEMPTY + st));
        tc_commit(bar(BAR_SFULL + u));
      }
      __syncwarp();
      kd += K_BYTES >> 4;
      if (++st == KS) { st = 0; ph ^= 1; kd = kd_base; }
    }
  } else if (warp >= 4) {
    // ---- softmax groups
    const int g = (warp - 4) >> 2;                  // group: tiles t = g, g + 2, ...
    const int q4 = warp & 3;
    const int row = q4 * 32 + lane;                 // TMEM lane == query row of the tile
    const uint32_t lane_addr = (uint32_t)(q4 * 32) << 16;
    float* mref = reinterpret_cast<float*>(sm + OFF_MREF);
    const uint32_t my_mref_bar = bar(BAR_MREF + g * 4 + q4), peer_mref_bar = bar(BAR_MREF + (g ^ 1) * 4 + q4);
    float m_loc = -INFINITY;
    int nwait = 0;
    for (int t = g; t < T; t += 2) {
      const int u = t & 3;
      mbar_wait(bar(BAR_SFULL + u), (t >> 2) & 1);
      tc_fence_after();
      const uint32_t s_addr = tmem + lane_addr + u * TK;
      uint32_t a[32], b[32];
      tmem_ld32_async(s_addr, a); tmem_ld32_async(s_addr + 32, b);
      tmem_ld_wait64(a, b);
      float m0 = __uint_as_float(a[0]), m1 = __uint_as_float(b[0]);
#pragma unroll
      for (int e = 1; e < 32; ++e) { m0 = fmaxf(m0, __uint_as_float(a[e])); m1 = fmaxf(m1, __uint_as_float(b[e])); }
      const float mt = fmaxf(m0, m1) * LOG2E;
      // take over the shared reference from the group that decided tile t-1
      if (t > 0) {
        mbar_wait(my_mref_bar, nwait & 1);
        ++nwait;
        const float m_sh = mref[row];
        m_loc = m_sh;
        const bool changed = mt > m_loc + RESCALE_TAU;
        if (__any_sync(0xffffffffu, changed)) {
          // raise the reference: rescale this row of O once P_{t-1} V_{t-1} (and everything before it) has landed
          const float m_new = changed ? mt + REF_MARGIN : m_loc;
          const float f = ex2(m_loc - m_new);
          mbar_wait(bar(BAR_PVDONE + ((t - 1) & 3)), ((t - 1) >> 2) & 1);
          tc_fence_after();
#pragma unroll 1
          for (int c = 0; c < CPAD; c += 32) {
            float ov[32];
            const uint32_t ta = tmem + lane_addr + COL_O + c;
            tmem_ld32(ta, ov);
#pragma unroll
            for (int e = 0; e < 32; ++e) ov[e] *= f;
            tmem_st32(ta, ov);
          }
          tmem_wait_st();
          m_loc = m_new;
        }
      } else {
        m_loc = mt + REF_MARGIN;
      }
      mref[row] = m_loc;
      if (t + 1 < T) {
        __syncwarp();
        if (lane == 0) mbar_arrive(peer_mref_bar);
      }
      // P = exp2(S*log2e - m_ref) in bf16 (two keys per 32-bit TMEM column).  The row sum is NOT accumulated here: V carries a
      // channel of ones (column C), so l = sum of the ROUNDED P comes out of the P.V product itself -- numerator and denominator
      // of the softmax see the same bf16 weights, and 64 additions per tile leave the softmax threads' instruction stream
      uint32_t packed[32];
#pragma unroll
      for (int e = 0; e < 32; e += 2) {
        const float p0 = ex2(fmaf(__uint_as_float(a[e]), LOG2E, -m_loc)), p1 = ex2(fmaf(__uint_as_float(a[e + 1]), LOG2E, -m_loc));
        // one exponential in four on the FMA pipe (the exp units are the busiest pipe of this loop: 54 % vs 10 % FMA); measured
        // 835 -> 892 TFLOP/s; 31 % polynomial: 880, 50 %: 850
        const float p2 = ex2(fmaf(__uint_as_float(b[e]), LOG2E, -m_loc)), p3 = ex2_poly(fmaf(__uint_as_float(b[e + 1]), LOG2E, -m_loc));
        __nv_bfloat162 h01 = __floats2bfloat162_rn(p0, p1), h23 = __floats2bfloat162_rn(p2, p3);
        packed[e >> 1] = *reinterpret_cast<uint32_t*>(&h01);
        packed[16 + (e >> 1)] = *reinterpret_cast<uint32_t*>(&h23);
      }
      tmem_st32_u(s_addr, packed);                       // P columns [0,32) of S[u]: every S column has been read
      tmem_wait_st();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar(BAR_PFULL + u));
    }
    // ---- epilogue: o = O / l, y = gamma*o + x, lse
    asm volatile("bar.sync 1, 256;" ::: "memory");           // the last reference has been published
    const float m_fin = (((T - 1) & 1) == g) ? m_loc : mref[row];   // decided by the group that owned the last tile
    mbar_wait(bar(BAR_PVDONE + ((T - 1) & 3)), ((T - 1) >> 2) & 1);
    tc_fence_after();
    const float ltot = tmem_ld1(tmem + lane_addr + COL_O + p.C);    // the ones channel of V: sum of the bf16 softmax weights of this row
    const float inv = 1.f / ltot;
    const size_t row0 = (size_t)sample * p.N + (size_t)qtile * TQ;
    if (g == 0) p.lse[row0 + row] = (m_fin + log2f(ltot)) * LN2;
    float* stg = reinterpret_cast<float*>(sm + OFF_V);       // the V ring is drained: all P.V products have completed
#pragma unroll 1
    for (int c = 0; c < CPAD / 2; c += 32) {
      float ov[32];
      const int c0 = g * (CPAD / 2) + c;
      tmem_ld32(tmem + lane_addr + COL_O + c0, ov);
#pragma unroll
      for (int e = 0; e < 32; e += 4)
        *reinterpret_cast<float4*>(stg + row * STG_STRIDE + c0 + e) = make_float4(ov[e] * inv, ov[e + 1] * inv, ov[e + 2] * inv, ov[e + 3] * inv);
    }
    tc_fence_before();
    asm volatile("bar.sync 1, 256;" ::: "memory");
    const float gm = __ldg(p.gamma);
    const int c4n = p.C >> 2;                                 // C is a multiple of 4 on this path (checked on the host)
    const int total = TQ * c4n;
    for (int idx = threadIdx.x - 128; idx < total; idx += 256) {
      const int r = idx / c4n, ch = (idx - r * c4n) << 2;
      const float4 on = *reinterpret_cast<const float4*>(stg + r * STG_STRIDE + ch);
      const size_t grow = row0 + r;
      const float4 xv = __ldg(reinterpret_cast<const float4*>(p.x + grow * p.x_pitch + ch));
      *reinterpret_cast<float4*>(p.o + grow * p.C + ch) = on;
      const float4 yv = make_float4(fmaf(gm, on.x, xv.x), fmaf(gm, on.y, xv.y), fmaf(gm, on.z, xv.z), fmaf(gm, on.w, xv.w));
      if (Y16) {
        const __nv_bfloat162 h0 = __floats2bfloat162_rn(yv.x, yv.y), h1 = __floats2bfloat162_rn(yv.z, yv.w);
        *reinterpret_cast<uint2*>(p.y16 + grow * p.y16_pitch + ch) = make_uint2(*reinterpret_cast<const uint32_t*>(&h0), *reinterpret_cast<const uint32_t*>(&h1));
      } else {
        *reinterpret_cast<float4*>(p.y + grow * p.y_pitch + ch) = yv;
      }
    }
  }
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u) : "memory");
  }
}

// fp32 -> fp16, round to nearest, saturating at the largest finite value (the network's q, k, v are O(10); an overflow must not become inf)
__device__ __forceinline__ __half f2h_sat(float f) { return __float2half_rn(fminf(fmaxf(f, -65504.f), 65504.f)); }

// ---- operand packing: fp32 NHWC slices -> tensor-core operands.  One warp per row: lanes 0-3 write the four 16-byte chunks of
// Q[row][0..32) (fp16), lanes 4-7 those of K (fp16), lanes 8-31 the 24 chunks of V[row][0..192) (bf16; zero beyond C except the
// channel of ones at column C, which makes the P.V product deliver the softmax denominator).
// SPLIT: Q and K rows are 64 halves wide: [hi(32) | lo(32)] with hi = fp16(v), lo = fp16(v - hi).
template <int SPLIT>
__global__ void __launch_bounds__(256) pack_qkv_kernel(const float* __restrict__ q, const float* __restrict__ k, int qk_pitch, int d, const float* __restrict__ v, int v_pitch,
                                                       int C, __half* __restrict__ Qh, __half* __restrict__ Kh, __nv_bfloat16* __restrict__ Vb, long long rows, int vec4) {
  const int lane = threadIdx.x & 31;
  const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
  constexpr int QK_ROW = DPAD * (1 + SPLIT);
  for (long long r = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5; r < rows; r += nwarps) {
    if (lane < 8) {
      const float* src = lane < 4 ? q : k;
      __half* dst = lane < 4 ? Qh : Kh;
      const int c0 = (lane & 3) * 8;
      __align__(16) __half h[8];
      __align__(16) __half l[8];
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        const float f = c0 + e < d ? __ldg(src + (size_t)r * qk_pitch + c0 + e) : 0.f;
        h[e] = __float2half_rn(f);
        if (SPLIT) l[e] = __float2half_rn(f - __half2float(h[e]));
      }
      *reinterpret_cast<uint4*>(dst + (size_t)r * QK_ROW + c0) = *reinterpret_cast<const uint4*>(h);
      if (SPLIT) *reinterpret_cast<uint4*>(dst + (size_t)r * QK_ROW + DPAD + c0) = *reinterpret_cast<const uint4*>(l);
    } else if (v != nullptr) {        // v == nullptr: the caller supplies the packed V operand (gdn_pam_fwd_args.v16)
      const int c0 = (lane - 8) * 8;
      __align__(16) __nv_bfloat16 h[8];
      if (vec4) {     // C % 4 == 0, 16-byte aligned rows: a group of four channels is entirely inside or entirely outside the C valid ones
#pragma unroll
        for (int g = 0; g < 2; ++g) {
          const int c = c0 + 4 * g;
          float4 f = c < C ? __ldg(reinterpret_cast<const float4*>(v + (size_t)r * v_pitch + c)) : make_float4(c == C ? 1.f : 0.f, 0.f, 0.f, 0.f);
          h[4 * g + 0] = __float2bfloat16_rn(f.x); h[4 * g + 1] = __float2bfloat16_rn(f.y);
          h[4 * g + 2] = __float2bfloat16_rn(f.z); h[4 * g + 3] = __float2bfloat16_rn(f.w);
        }
      } else {
#pragma unroll
        for (int e = 0; e < 8; ++e) h[e] = __float2bfloat16_rn(c0 + e < C ? __ldg(v + (size_t)r * v_pitch + c0 + e) : (c0 + e == C ? 1.f : 0.f));
      }
      *reinterpret_cast<uint4*>(Vb + (size_t)r * CPAD + c0) = *reinterpret_cast<const uint4*>(h);
    }
  }
}
// aug == nullptr: columns d, d+1 = 1 (the K side); else columns d, d+1 = -aug[row] split into fp16 hi + lo (the Q side: -lse), so that the
// logit product of the backward kernels comes out of the tensor core as S - lse (exact to 2^-22 |lse|) and no thread ever loads lse.
// split: rows are 64 halves wide, [hi(32) | lo(32)]; the lo chunk holds fp16(v - hi) in columns [0, d) and zeros elsewhere (the augmentation
// columns live in the hi chunk only, so the two cross products add nothing to them).
__global__ void __launch_bounds__(256) pack_qk_kernel(const float* __restrict__ src, int pitch, int d, __half* __restrict__ dst, long long rows,
                                                      const float* __restrict__ aug, int augment, int split) {
  const long long total = rows * DPAD;
  const int row_w = split ? 2 * DPAD : DPAD;
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
    long long r = idx / DPAD; int c = (int)(idx % DPAD);
    float v = c < d ? src[(size_t)r * pitch + c] : 0.f;
    if (augment && (c == d || c == d + 1)) {
      if (!aug) v = 1.f;
      else {
        const float a = -__ldg(aug + r);
        const float hi = __half2float(__float2half_rn(a));
        v = c == d ? hi : a - hi;
      }
    }
    const __half h = __float2half_rn(v);
    dst[(size_t)r * row_w + c] = h;
    if (split) dst[(size_t)r * row_w + DPAD + c] = c < d ? __float2half_rn(v - __half2float(h)) : __float2half_rn(0.f);
  }
}

// 2-D fp16 row-major tensor [rows][cols]; box = [box_rows][box_cols]
static int make_map(CUtensorMap* m, const void* ptr, uint64_t rows, uint64_t cols, uint32_t box_rows, uint32_t box_cols, CUtensorMapSwizzle sw,
                    CUtensorMapDataType dt = CU_TENSOR_MAP_DATA_TYPE_FLOAT16) {
  EncodeTiledFn enc = get_encode();
  if (!enc) { set_error("pam_tc: cuTensorMapEncodeTiled unavailable"); return GDN_ECUDA; }
  cuuint64_t gdim[2] = {cols, rows};
  cuuint64_t gstride[1] = {cols * 2};
  cuuint32_t box[2] = {box_cols, box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(m, dt, 2, const_cast<void*>(ptr), gdim, gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, sw,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { set_error("pam_tc: cuTensorMapEncodeTiled failed (%d)", (int)r); return GDN_ECUDA; }
  return GDN_OK;
}
static inline size_t align256(size_t v) { return (v + 255) & ~(size_t)255; }
}  // namespace pamtc
}  // namespace gdn

using namespace gdn;
using namespace gdn::pamtc;

extern "C" int gdn_pam_tc_bwd_init(void);
extern "C" int gdn_pam_tc_init(void) {
  GDN_CHECK_CUDA(cudaFuncSetAttribute(pam_flash_fwd_kernel<0, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, FwdSmem<0>::SMEM_BYTES));
  GDN_CHECK_CUDA(cudaFuncSetAttribute(pam_flash_fwd_kernel<1, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, FwdSmem<0>::SMEM_BYTES));
  GDN_CHECK_CUDA(cudaFuncSetAttribute(pam_flash_fwd_kernel<0, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, FwdSmem<1>::SMEM_BYTES));
  GDN_CHECK_CUDA(cudaFuncSetAttribute(pam_flash_fwd_kernel<1, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, FwdSmem<1>::SMEM_BYTES));
  return gdn_pam_tc_bwd_init();
}

static inline bool pam_tc_precision(int p) { return p == GDN_PREC_FP16 || p == GDN_PREC_FP16X3; }

extern "C" size_t gdn_pam_tc_fwd_ws_bytes(const gdn_pam_fwd_args* a) {
  size_t rows = (size_t)a->B * a->N;
  const size_t qk_row = DPAD * 2 * (a->precision == GDN_PREC_FP16X3 ? 2 : 1);
  return 2 * align256(rows * qk_row) + align256(rows * CPAD * 2);
}

extern "C" int gdn_pam_tc_fwd(const gdn_pam_fwd_args* a, gdn_stream_t s) {
  GDN_CHECK_ARG(pam_tc_precision(a->precision));
  const int split = a->precision == GDN_PREC_FP16X3;
  GDN_CHECK_ARG(a->N % TQ == 0 && a->d <= DPAD && a->C < CPAD && a->C % 4 == 0);      // C < CPAD: column C of V is the channel of ones
  GDN_CHECK_ARG(a->x_pitch % 4 == 0 && a->y_pitch % 4 == 0);
  GDN_CHECK_ARG(((uintptr_t)a->x & 15) == 0 && ((uintptr_t)a->y & 15) == 0 && ((uintptr_t)a->o & 15) == 0);
  if (!a->ws || a->ws_bytes < gdn_pam_tc_fwd_ws_bytes(a)) { set_error("gdn_pam_fwd(fp16): workspace too small"); return GDN_EWORKSPACE; }
  const size_t rows = (size_t)a->B * a->N;
  const size_t qk_cols = (size_t)DPAD * (split ? 2 : 1);
  char* w = reinterpret_cast<char*>(a->ws);
  __half* Qh = reinterpret_cast<__half*>(w); w += align256(rows * qk_cols * 2);
  __half* Kh = reinterpret_cast<__half*>(w); w += align256(rows * qk_cols * 2);
  const __nv_bfloat16* Vb = a->v16 ? reinterpret_cast<const __nv_bfloat16*>(a->v16) : reinterpret_cast<__nv_bfloat16*>(w);
  GDN_CHECK_ARG(((uintptr_t)Vb & 15) == 0);
  cudaStream_t st = as_stream(s);
  const long long pb = cdiv((long long)rows, 8);
  const unsigned pgrid = (unsigned)(pb < 16 * kNumSMs ? pb : 16 * kNumSMs);
  const int vec4 = a->C % 4 == 0 && a->v_pitch % 4 == 0 && ((uintptr_t)a->v & 15) == 0;
  __nv_bfloat16* vdst = a->v16 ? nullptr : reinterpret_cast<__nv_bfloat16*>(w);
  if (split) pack_qkv_kernel<1><<<pgrid, 256, 0, st>>>(a->q, a->k, a->qk_pitch, a->d, a->v16 ? nullptr : a->v, a->v_pitch, a->C, Qh, Kh, vdst, (long long)rows, vec4);
  else pack_qkv_kernel<0><<<pgrid, 256, 0, st>>>(a->q, a->k, a->qk_pitch, a->d, a->v16 ? nullptr : a->v, a->v_pitch, a->C, Qh, Kh, vdst, (long long)rows, vec4);
  GDN_CHECK_LAUNCH();
  CUtensorMap mq, mk, mv;
  int rc;
  if ((rc = make_map(&mq, Qh, rows, qk_cols, TQ, DPAD, CU_TENSOR_MAP_SWIZZLE_64B)) != GDN_OK) return rc;
  if ((rc = make_map(&mk, Kh, rows, qk_cols, TK, DPAD, CU_TENSOR_MAP_SWIZZLE_64B)) != GDN_OK) return rc;
  if ((rc = make_map(&mv, Vb, rows, CPAD, 64, 64, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16)) != GDN_OK) return rc;
  FwdParams p;
  p.x = a->x; p.x_pitch = a->x_pitch; p.gamma = a->gamma; p.o = a->o; p.y = a->y; p.y_pitch = a->y_pitch; p.lse = a->lse;
  p.B = a->B; p.N = a->N; p.C = a->C; p.tiles_per_sample = a->N / TQ;
  p.n_pv = (a->C + 1 + 15) & ~15;
  p.y16 = reinterpret_cast<__nv_bfloat16*>(a->y16); p.y16_pitch = a->y16_pitch;
  GDN_CHECK_ARG(a->y || a->y16);
  if (a->y16) GDN_CHECK_ARG(a->y16_pitch >= a->C && a->y16_pitch % 4 == 0 && ((uintptr_t)a->y16 & 7) == 0);
  const int grid = a->B * p.tiles_per_sample;
  if (a->y16 && !a->y) {
    if (split) pam_flash_fwd_kernel<1, 1><<<grid, NTHREADS, FwdSmem<1>::SMEM_BYTES, st>>>(mq, mk, mv, p);
    else pam_flash_fwd_kernel<1, 0><<<grid, NTHREADS, FwdSmem<0>::SMEM_BYTES, st>>>(mq, mk, mv, p);
  } else {
    GDN_CHECK_ARG(a->y && !a->y16);      // the bf16 block replaces the fp32 output (both at once is not a supported combination)
    if (split) pam_flash_fwd_kernel<0, 1><<<grid, NTHREADS, FwdSmem<1>::SMEM_BYTES, st>>>(mq, mk, mv, p);
    else pam_flash_fwd_kernel<0, 0><<<grid, NTHREADS, FwdSmem<0>::SMEM_BYTES, st>>>(mq, mk, mv, p);
  }
  GDN_CHECK_LAUNCH();
  return GDN_OK;
}

// =====================================================================================================================
// Fused backward of the position-attention core (autograd of generator.py:115-122; formulas: SURVEY appendix C).
// Two launches of ONE kernel template, both flash-style (the N x N maps P, dP, dS never leave the SM):
//   MODE 0 (dQ)    : CTA = one sample x 128 query rows, loops over 64-key blocks
//                    X = Q_i K_j^T, Y = dy_i V_j^T, P = exp(X - L_i), dS = P*(Y - rowdot_i), dQ_i += dS K_j
//   MODE 1 (dK,dV) : CTA = one sample x 128 key rows, loops over 64-query blocks (everything transposed, so that the TMEM
//                    lane is the key): X = K_j Q_i^T, Y = V_j dy_i^T, P^T, dS^T, dK_j += dS^T Q_i, dV_j += P^T dy_i
//                    (dy_i is consumed from ONE shared-memory tile both as K-major B operand of Y and as MN-major B
//                    operand of dV -- no transposed copy of dy exists).
// gamma is folded into the epilogue (dO = gamma*dy): dQ, dK, dV are scaled by gamma on the way out.
// Each CTA owns its output rows: no atomics, bitwise deterministic.
// lse and rowdot never reach the threads: the Q rows carry -lse (fp16 hi + lo) and the K rows 1 in two spare columns of the padded logit
// operands, the dy rows carry -rowdot (bf16 hi + lo) and the V rows 1 in two spare channel columns, so the tensor core delivers
// X = S - lse and Y = dP - rowdot directly (ncu before: the per-column lse / rowdot broadcasts of the dK/dV launch were 39 % of the
// shared-memory data pipe next to 57 % tensor-core operand reads).
// Operand precision: every operand is fp16 (11-bit mantissa), fp32 accumulation.  Logits fp16 x fp16 (hi + lo split in GDN_PREC_FP16X3)
// exactly as in the forward kernel, so P matches the saved log-sum-exp.  The gradient-carrying operands fit fp16's range through ONE
// power-of-two scale per call: sc = 2^-ceil(log2 max|dy|) (device scalar, computed by the rowdot pass that reads dy anyway), dy*sc in
// [-1, 1]; dS = P*(dP - rowdot)*sc is bounded by 2*C*max|v| and saturates instead of overflowing; P <= 1.  1/sc is folded into the
// epilogue with gamma.  With bf16 gradient operands (round 1) the cancellation in dP - rowdot amplified the 2^-9 operand rounding to
// 7e-3 on dq / dk at reference-scale logits (std 10); fp16: 1.4e-3 (tools/measure_pam_split.py, float64 emulation in DESIGN.md 4).
//   warp 0: TMA producer   warp 1: tcgen05.mma issuer   warp 2: TMEM allocator   warps 4-11: softmax/dS (thread = TMEM lane = row)
// TMEM: X0 X1 [0,128)  Y0 Y1 [128,256)  small accumulator (dQ | dK) [256,288)  dV [288,480)
namespace gdn {
namespace pamtc {
namespace bwd {
constexpr int TO = 128, TI = 64, NCH = CPAD / 64, ST = 4;
constexpr int OQK_CHUNK = TO * DPAD * 2;              // 8 KB  fp16 [128][32], SWIZZLE_64B
constexpr int OC_CHUNK = TO * 128;                    // 16 KB bf16 [128][64], SWIZZLE_128B
constexpr int IQK_CHUNK = TI * DPAD * 2;              // 4 KB
constexpr int IT_BYTES = DPAD * TI * 2;               // 4 KB  bf16 [32][64]
constexpr int IC_CHUNK = TI * 128;                    // 8 KB  bf16 [64][64]
// SPLIT = 1 (GDN_PREC_FP16X3): the logit operands are fp16 hi + lo pairs, two 32-column chunks per tile, exactly as in the forward kernel
// (X = o_lo i_hi^T + o_hi i_lo^T + o_hi i_hi^T; the -lse / 1 augmentation columns live in the hi chunks only).
template <int SPLIT>
struct BwdSmem {
  static constexpr int OQK_BYTES = OQK_CHUNK * (1 + SPLIT);
  static constexpr int IQK_BYTES = IQK_CHUNK * (1 + SPLIT);
  static constexpr int OFF_OQK = 0;
  static constexpr int OFF_OC = OFF_OQK + OQK_BYTES;           // 8192 / 16384
  static constexpr int OFF_IN = OFF_OC + NCH * OC_CHUNK;       // 57344 / 65536
  static constexpr int IN_IQK = 0, IN_IT = IQK_BYTES, IN_IC = IN_IT + IT_BYTES;   // IN_IC: 8192 / 12288
  static constexpr int IN_BYTES = IN_IC + NCH * IC_CHUNK + 1024;           // 33 / 37 KB per stage
  static constexpr int OFF_BARS = OFF_IN + ST * IN_BYTES;      // (P and dS never touch shared memory: they go back to TMEM over the consumed X / Y columns)
  static constexpr int OFF_TSLOT = OFF_BARS + 256;
  static constexpr int SMEM = OFF_TSLOT + 16 + 1024;
  static_assert(OFF_OC % 1024 == 0 && OFF_IN % 1024 == 0 && IN_IC % 1024 == 0 && IN_BYTES % 1024 == 0, "swizzle atoms");
  static_assert(SMEM <= 227 * 1024, "shared memory budget");
};
enum { B_OUT = 0, B_INFULL = 1, B_INEMPTY = B_INFULL + ST, B_XFULL = B_INEMPTY + ST, B_YFULL = B_XFULL + 2, B_DSFULL = B_YFULL + 2, B_ACC = B_DSFULL + 2,
       B_OCT = B_ACC + 1, B_COUNT = B_OCT + 1 };
static_assert(B_COUNT * 8 <= 256, "barrier area");
// TMEM columns.  MODE 0 (dQ):    X0 X1 [0,128)  Y0 Y1 [128,256)  dQ [256,288)  dy_i (A operand of Y, loop invariant) [288,384)  Q_i hi|lo [384,416)
//               MODE 1 (dK,dV): X0 X1 [0,128)  Y0 Y1 [128,256)  dK [256,288)  dV [288,480)                                     K_j hi|lo [480,512)
// The loop-invariant outer operands live in TMEM (copied once per CTA by the softmax warps): an A operand in shared memory is re-read by every
// MMA of every inner block (4 KB per K = 16 step), and the dK/dV launch is bound by the shared-memory pipe (ncu round 1: 57 % tensor-core operand
// reads + 39 % LSU).  MODE 1 has no room for V_j (96 columns) next to two Y buffers: tried with ONE Y buffer in round 2 -- correct, but the
// serialisation Y_{t+1} after dS_t cost more (2.76 -> 3.44 ms per backward) than the operand reads it saved.
constexpr uint32_t COL_X = 0, COL_Y = 128, COL_SMALL = 256, COL_BIG = 288;
template <int MODE> struct BwdCols { static constexpr uint32_t QK = MODE == 0 ? 384 : 480; };
#ifndef GDN_PAM_BWD_POLY
#define GDN_PAM_BWD_POLY 0
#endif
constexpr bool BWD_POLY = GDN_PAM_BWD_POLY != 0;      // one exponential in four on the FMA pipe as in the forward kernel: measured SLOWER here in both rounds
                                                       // (round 2, TMEM-resident operands: 2.99 -> 3.04 ms per backward) -- the exp units are not this kernel's limiter


// kind::f16 instruction descriptor for bf16 operands (see conv_tc.cu): b_mn = 1 makes B MN-major
__host__ __device__ constexpr uint32_t idesc_bf16(int M, int N, int b_mn) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)b_mn << 16) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ uint64_t smem_desc_lbo(uint32_t addr, uint32_t lbo_bytes, uint32_t sbo_bytes, uint32_t layout_type) {
  return (uint64_t)((addr >> 4) & 0x3FFF) | ((uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16) | ((uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32) | ((uint64_t)1 << 46) |
         ((uint64_t)layout_type << 61);
}
__device__ __forceinline__ void bulk_load(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}

// two fp32 -> packed fp16x2, round to nearest, saturating to +-65504 (an overflowing dS must not become inf: inf * 0 = NaN in the next MMA)
__device__ __forceinline__ uint32_t pack_half2_sat(float lo, float hi) {
  uint32_t r;
  asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
  return r;
}

struct Params {
  const float* lse; const float* rowdot; const float* gamma; const float* scale;   // scale[1] = sc (power of two applied to dy and rowdot), scale[2] = 1 / sc
  float* dq; float* dk; float* dv;
  int B, N, C, d;
};


template <int MODE, int SPLIT>
__global__ void __launch_bounds__(NTHREADS, 1)
pam_flash_bwd_kernel(const __grid_constant__ CUtensorMap mapQh, const __grid_constant__ CUtensorMap mapKh, const __grid_constant__ CUtensorMap mapV,
                     const __grid_constant__ CUtensorMap mapDY, const __grid_constant__ CUtensorMap mapQt, const __grid_constant__ CUtensorMap mapKt, const Params p) {
  using SM = BwdSmem<SPLIT>;
  constexpr uint32_t COL_QK = BwdCols<MODE>::QK;          // outer logit operand (hi | lo chunks, 16 columns each)
  constexpr int OQK_BYTES = SM::OQK_BYTES, IQK_BYTES = SM::IQK_BYTES, OFF_OQK = SM::OFF_OQK, OFF_OC = SM::OFF_OC, OFF_IN = SM::OFF_IN, IN_IQK = SM::IN_IQK,
                IN_IT = SM::IN_IT, IN_IC = SM::IN_IC, IN_BYTES = SM::IN_BYTES, OFF_BARS = SM::OFF_BARS, OFF_TSLOT = SM::OFF_TSLOT;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* sm = smem_raw + (base - smem_u32(smem_raw));
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int blocks_per_sample = p.N / TO;
  const int sample = blockIdx.x / blocks_per_sample, oblk = blockIdx.x % blocks_per_sample;
  const int T = p.N / TI;
  const int row0 = sample * p.N + oblk * TO;           // first outer row in the flattened [B*N] row space
  auto bar = [&](int i) { return base + OFF_BARS + 8 * i; };
  volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(sm + OFF_TSLOT);
  const CUtensorMap* m_oqk = MODE == 0 ? &mapQh : &mapKh;
  const CUtensorMap* m_iqk = MODE == 0 ? &mapKh : &mapQh;
  const CUtensorMap* m_oc = MODE == 0 ? &mapDY : &mapV;
  const CUtensorMap* m_ic = MODE == 0 ? &mapV : &mapDY;
  const CUtensorMap* m_it = MODE == 0 ? &mapKt : &mapQt;

  if (threadIdx.x == 0) {
    mbar_init(bar(B_OUT), 1);
    for (int i = 0; i < ST; ++i) { mbar_init(bar(B_INFULL + i), 1); mbar_init(bar(B_INEMPTY + i), 1); }
    for (int i = 0; i < 2; ++i) {
      mbar_init(bar(B_XFULL + i), 1); mbar_init(bar(B_YFULL + i), 1);
      mbar_init(bar(B_DSFULL + i), 4);                                         // 4 warps: buffer i belongs to softmax warp group i
    }
    mbar_init(bar(B_ACC), 1);
    mbar_init(bar(B_OCT), 8);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(base + OFF_TSLOT), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;

  if (warp == 0) {
    if (lane == 0) {   // ---- TMA producer
      mbar_expect_tx(bar(B_OUT), OQK_BYTES + NCH * OC_CHUNK);
      tma_load_2d(base + OFF_OQK, m_oqk, bar(B_OUT), 0, row0);
      tma_load_2d(base + OFF_OQK + OQK_CHUNK / 2, m_oqk, bar(B_OUT), 0, row0 + 64);
      if (SPLIT) {
        tma_load_2d(base + OFF_OQK + OQK_CHUNK, m_oqk, bar(B_OUT), DPAD, row0);
        tma_load_2d(base + OFF_OQK + OQK_CHUNK + OQK_CHUNK / 2, m_oqk, bar(B_OUT), DPAD, row0 + 64);
      }
      for (int c = 0; c < NCH; ++c) {
        tma_load_2d(base + OFF_OC + c * OC_CHUNK, m_oc, bar(B_OUT), c * 64, row0);
        tma_load_2d(base + OFF_OC + c * OC_CHUNK + OC_CHUNK / 2, m_oc, bar(B_OUT), c * 64, row0 + 64);
      }
      for (int t = 0; t < T; ++t) {
        const int s = t % ST;
        if (t >= ST) mbar_wait(bar(B_INEMPTY + s), ((t / ST) - 1) & 1);
        const uint32_t st = base + OFF_IN + s * IN_BYTES;
        const int irow = sample * p.N + t * TI;
        mbar_expect_tx(bar(B_INFULL + s), IQK_BYTES + IT_BYTES + NCH * IC_CHUNK);
        tma_load_2d(st + IN_IQK, m_iqk, bar(B_INFULL + s), 0, irow);
        if (SPLIT) tma_load_2d(st + IN_IQK + IQK_CHUNK, m_iqk, bar(B_INFULL + s), DPAD, irow);
        tma_load_2d(st + IN_IT, m_it, bar(B_INFULL + s), t * TI, sample * DPAD);
        for (int c = 0; c < NCH; ++c) tma_load_2d(st + IN_IC + c * IC_CHUNK, m_ic, bar(B_INFULL + s), c * 64, irow);
      }
    }
  } else if (warp == 1) {
    // ---- MMA issuer: warp-uniform loop, one elected lane issues; descriptors are base descriptors plus 16-byte-unit offsets
    constexpr uint32_t ID_X = idesc_f16(TO, TI), ID_Y = idesc_f16(TO, TI), ID_S = idesc_f16(TO, DPAD), ID_B = idesc_f16(TO, CPAD) | (1u << 16);   // all fp16; ID_B: B MN-major
    const uint64_t d64 = smem_desc(base, 512, LAYOUT_SW64), d128 = smem_desc(base, 1024, LAYOUT_SW128);
    const uint64_t dmn = smem_desc_lbo(base, IC_CHUNK, 1024, LAYOUT_SW128);     // MN-major: 64-channel groups 8 KB apart, 8-row groups 1 KB apart
    const uint64_t oqk_d = d64 + (OFF_OQK >> 4), oc_d = d128 + (OFF_OC >> 4);
    auto issue_x = [&](int t) {
      const int b = t & 1, s = t % ST;
      mbar_wait_spin(bar(B_INFULL + s), (uint32_t)(t / ST) & 1u);
      // X[b] / Y[b] hold P_{t-2} / dS_{t-2} until the accumulating products of block t-2 have read them; those were issued before
      // this call (same thread, same pipe: issue order), after the softmax group's DSFULL hand-over -- no separate "free" barriers
      tc_fence_after();
      const uint32_t st_off = (uint32_t)(OFF_IN + s * IN_BYTES) >> 4;
      const uint64_t iqk_d = d64 + st_off + (IN_IQK >> 4);
      const uint32_t xd = tmem + COL_X + b * TI, ahi = tmem + COL_QK, alo = tmem + COL_QK + 16;     // A = outer logit operand from TMEM: K = 16 per 8 columns
      if (elect_one()) {
        if (SPLIT) {      // same order of terms as the forward kernel
          umma_f16_ts_i<0>(xd, alo, iqk_d, ID_X);                                    // o_lo i_hi^T
          umma_f16_ts_i<1>(xd, alo + 8, iqk_d + 2, ID_X);
          umma_f16_ts_i<1>(xd, ahi, iqk_d + (IQK_CHUNK >> 4), ID_X);                 // o_hi i_lo^T
          umma_f16_ts_i<1>(xd, ahi + 8, iqk_d + (IQK_CHUNK >> 4) + 2, ID_X);
          umma_f16_ts_i<1>(xd, ahi, iqk_d, ID_X);                                    // o_hi i_hi^T (carries -lse)
        } else {
          umma_f16_ts_i<0>(xd, ahi, iqk_d, ID_X);
        }
        umma_f16_ts_i<1>(xd, ahi + 8, iqk_d + 2, ID_X);
        tc_commit(bar(B_XFULL + b));
      }
      __syncwarp();
    };
    auto issue_y = [&](int t) {
      const int b = t & 1, s = t % ST;
      const uint32_t ycol = tmem + COL_Y + b * TI;
      const uint32_t st_off = (uint32_t)(OFF_IN + s * IN_BYTES) >> 4;          // (INFULL of this block has been awaited by issue_x)
      const uint64_t ic_d = d128 + st_off + (IN_IC >> 4);
      if (elect_one()) {
#pragma unroll
        for (int c = 0; c < NCH; ++c)
#pragma unroll
          for (int ks = 0; ks < 4; ++ks) {
            const uint64_t ad = oc_d + (uint64_t)(c * (OC_CHUNK >> 4) + ks * 2), bd = ic_d + (uint64_t)(c * (IC_CHUNK >> 4) + ks * 2);
            if (MODE == 0) {      // A = dy_i from TMEM (copied there once per CTA): 16 channels = 8 columns per step
              const uint32_t at = tmem + COL_BIG + (c * 4 + ks) * 8;
              if (c == 0 && ks == 0) umma_f16_ts_i<0>(ycol, at, bd, ID_Y); else umma_f16_ts_i<1>(ycol, at, bd, ID_Y);
            } else {
              if (c == 0 && ks == 0) umma_f16_i<0>(ycol, ad, bd, ID_Y); else umma_f16_i<1>(ycol, ad, bd, ID_Y);
            }
          }
        tc_commit(bar(B_YFULL + b));
      }
      __syncwarp();
    };
    auto issue_acc = [&](int t) {
      const int b = t & 1, s = t % ST;
      mbar_wait_spin(bar(B_DSFULL + b), (t >> 1) & 1);
      tc_fence_after();
      const uint32_t st_off = (uint32_t)(OFF_IN + s * IN_BYTES) >> 4;
      const uint64_t it_d = d128 + st_off + (IN_IT >> 4);
      if (elect_one()) {
        // small accumulator += dS x inner^T tile [32][64]; A = dS from TMEM, written by the owning softmax warp group over the consumed
        // Y[b] columns: 64 inner rows = columns [0, 32), 16 per step
        const uint32_t ya = tmem + COL_Y + b * TI;
        umma_f16_ts(tmem + COL_SMALL, ya, it_d, ID_S, t > 0 ? 1u : 0u);
        umma_f16_ts_i<1>(tmem + COL_SMALL, ya + 8, it_d + 2, ID_S);
        umma_f16_ts_i<1>(tmem + COL_SMALL, ya + 16, it_d + 4, ID_S);
        umma_f16_ts_i<1>(tmem + COL_SMALL, ya + 24, it_d + 6, ID_S);
        if (MODE == 1) {
          // dV += P^T x dy_i: A = P^T from TMEM (over the consumed X[b] columns), B = the [64 q][192 ch] dy tile read MN-major (2 KB per step)
          const uint32_t xa = tmem + COL_X + b * TI;
          const uint64_t icmn_d = dmn + st_off + (IN_IC >> 4);
          umma_f16_ts(tmem + COL_BIG, xa, icmn_d, ID_B, t > 0 ? 1u : 0u);
          umma_f16_ts_i<1>(tmem + COL_BIG, xa + 8, icmn_d + 128, ID_B);
          umma_f16_ts_i<1>(tmem + COL_BIG, xa + 16, icmn_d + 256, ID_B);
          umma_f16_ts_i<1>(tmem + COL_BIG, xa + 24, icmn_d + 384, ID_B);
        }
        tc_commit(bar(B_INEMPTY + s));
      }
      __syncwarp();
    };
    mbar_wait_spin(bar(B_OUT), 0);
    mbar_wait_spin(bar(B_OCT), 0);           // the loop-invariant outer operands have been copied into TMEM by the softmax warps
    // two X and two Y buffers: block t+1's logits and dP are formed while the softmax groups work on block t
    issue_x(0); issue_y(0);
    for (int t = 0; t < T; ++t) {
      if (t + 1 < T) { issue_x(t + 1); issue_y(t + 1); }
      issue_acc(t);
    }
    if (elect_one()) tc_commit(bar(B_ACC));
    __syncwarp();
  } else if (warp >= 4) {
    const int wg = (warp - 4) >> 2;                  // warp group: owns the inner blocks (and X/Y buffers) of parity wg
    const int q4 = warp & 3;
    const int row = q4 * 32 + lane;                  // TMEM lane = outer row
    const uint32_t lane_addr = (uint32_t)(q4 * 32) << 16;
    {
      mbar_wait(bar(B_OUT), 0);
      // The outer logit operand (Q_i in MODE 0, K_j in MODE 1; hi and lo chunks) goes to TMEM once: the A operand of every X MMA.  Thread = row; a
      // 64-byte SWIZZLE_64B row holds four 16-byte pieces, piece p at physical position p ^ ((row >> 1) & 3); 8 halves = 4 TMEM columns per piece.
      // SPLIT: warp group 0 copies the hi chunk (columns [0,16)), group 1 the lo chunk ([16,32)); else each group copies two pieces of the one chunk.
      {
        const uint8_t* src = sm + OFF_OQK + (SPLIT ? wg * OQK_CHUNK : 0) + (row >> 3) * 512 + (row & 7) * 64;
        const int sw = (row >> 1) & 3;
        if (SPLIT) {
          uint32_t v[16];
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const uint4 q = *reinterpret_cast<const uint4*>(src + ((j ^ sw) << 4));
            v[4 * j] = q.x; v[4 * j + 1] = q.y; v[4 * j + 2] = q.z; v[4 * j + 3] = q.w;
          }
          tmem_st16_u(tmem + lane_addr + COL_QK + wg * 16, v);
        } else {
          uint32_t v[8];
#pragma unroll
          for (int j = 0; j < 2; ++j) {
            const uint4 q = *reinterpret_cast<const uint4*>(src + (((wg * 2 + j) ^ sw) << 4));
            v[4 * j] = q.x; v[4 * j + 1] = q.y; v[4 * j + 2] = q.z; v[4 * j + 3] = q.w;
          }
          tmem_st8_u(tmem + lane_addr + COL_QK + wg * 8, v);
        }
      }
      if (MODE == 0) {
        // dy_i (outer tile, loop invariant, A operand of Y = dy_i V_j^T) goes to TMEM too: with A in shared memory every one of the
        // 12 MMAs of a block re-reads its 4 KB A slice (49 instead of 33 cycles each, and the shared-memory pipe is the kernel's bound).
        // Warp group w copies 16-byte pieces 4w..4w+3 of each 128-byte swizzled row chunk: 8 values = 4 TMEM columns per piece.
#pragma unroll
        for (int c = 0; c < NCH; ++c) {
          const uint8_t* src = sm + OFF_OC + c * OC_CHUNK + (row >> 3) * 1024 + (row & 7) * 128;
          uint32_t v[16];
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const uint4 q = *reinterpret_cast<const uint4*>(src + (((wg * 4 + j) ^ (row & 7)) << 4));
            v[4 * j] = q.x; v[4 * j + 1] = q.y; v[4 * j + 2] = q.z; v[4 * j + 3] = q.w;
          }
          tmem_st16_u(tmem + lane_addr + COL_BIG + c * 32 + wg * 16, v);
        }
      }
      tmem_wait_st();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar(B_OCT));
    }
    // The two warp groups ping-pong on alternate inner blocks (group w owns the X/Y buffers of parity w and all 64 columns of its
    // blocks, 32 at a time): while one group is in its MUFU-bound exp phase the other loads Y, forms dS and hands it to the MMA warp.
    for (int t = wg; t < T; t += 2) {
      const int b = wg, s = t % ST;
      mbar_wait(bar(B_XFULL + b), (t >> 1) & 1);
      tc_fence_after();
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const int col0 = h * 32;
        float x[32];
        tmem_ld32(tmem + lane_addr + COL_X + b * TI + col0, x);
        // X already is S - lse (the operands carry -lse and 1 in two spare columns of the logit product): P = exp(X) <= 1
#pragma unroll
        for (int i = 0; i < 32; ++i) x[i] = (BWD_POLY && (i & 3) == 3) ? ex2_poly(x[i] * LOG2E) : ex2(x[i] * LOG2E);
        if (h == 0) { mbar_wait(bar(B_YFULL + b), (t >> 1) & 1); tc_fence_after(); }
        float y[32];
        tmem_ld32(tmem + lane_addr + COL_Y + b * TI + col0, y);
        uint32_t ds_pk[16], p_pk[16];
        // Y already is dy V^T - rowdot (same trick: -rowdot and 1 in two spare channel columns): dS = P * Y
#pragma unroll
        for (int i = 0; i < 32; i += 2) {
          ds_pk[i >> 1] = pack_half2_sat(x[i] * y[i], x[i + 1] * y[i + 1]);
          if (MODE == 1) {
            __half2 pa = __floats2half2_rn(x[i], x[i + 1]);          // P <= 1
            p_pk[i >> 1] = *reinterpret_cast<uint32_t*>(&pa);
          }
        }
        // dS (and P) go back to TMEM over already consumed Y (X) columns -- inner rows [32h, 32h+32) -> columns [16h, 16h+16) -- as the A
        // operands of the accumulating products: no shared-memory round trip, no proxy fence
        tmem_st16_u(tmem + lane_addr + COL_Y + b * TI + h * 16, ds_pk);
        if (MODE == 1) tmem_st16_u(tmem + lane_addr + COL_X + b * TI + h * 16, p_pk);
      }
      tmem_wait_st();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar(B_DSFULL + b));
    }
    // ---- epilogue
    mbar_wait(bar(B_ACC), 0);
    tc_fence_after();
    const float g = __ldg(p.gamma) * __ldg(p.scale + 2);       // gamma / sc: an exact power-of-two rescale
    const size_t grow = (size_t)row0 + row;
    if (wg == 0) {
      float a[32];
      tmem_ld32(tmem + lane_addr + COL_SMALL, a);
      float* dst = (MODE == 0 ? p.dq : p.dk) + grow * p.d;
#pragma unroll
      for (int i = 0; i < 32; ++i)
        if (i < p.d) dst[i] = g * a[i];
    }
    if (MODE == 1) {
#pragma unroll 1
      for (int c = 0; c < CPAD / 2; c += 32) {
        float a[32];
        const int c0 = wg * (CPAD / 2) + c;
        tmem_ld32(tmem + lane_addr + COL_BIG + c0, a);
#pragma unroll
        for (int i = 0; i < 32; i += 4) {
          const int ch = c0 + i;
          if (ch < p.C) *reinterpret_cast<float4*>(p.dv + grow * p.C + ch) = make_float4(g * a[i], g * a[i + 1], g * a[i + 2], g * a[i + 3]);
        }
      }
    }
    tc_fence_before();
  }
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u) : "memory");
  }
}

// rowdot[m] = sum_c dy[m][c] * o[m][c] (one warp per row) and, on the way, max |dy| over the whole tensor (atomicMax on the bit pattern of a
// non-negative float; slot[0] must be zero before the launch)
__global__ void __launch_bounds__(256) rowdot_amax_kernel(const float* __restrict__ dy, int dy_pitch, const float* __restrict__ o, int o_pitch, long long M, int C,
                                                          float* __restrict__ out, unsigned* __restrict__ slot) {
  const int lane = threadIdx.x & 31;
  const long long warp = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5;
  const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
  float amax = 0.f;
  const bool v4 = (C % 4 == 0) && (dy_pitch % 4 == 0) && (o_pitch % 4 == 0) && ((reinterpret_cast<uintptr_t>(dy) | reinterpret_cast<uintptr_t>(o)) & 15) == 0;
  for (long long m = warp; m < M; m += nwarps) {
    float s = 0.f;
    if (v4) {
      for (int c = lane * 4; c < C; c += 128) {
        const float4 a = __ldg(reinterpret_cast<const float4*>(dy + (size_t)m * dy_pitch + c)), b = __ldg(reinterpret_cast<const float4*>(o + (size_t)m * o_pitch + c));
        s = fmaf(a.x, b.x, s); s = fmaf(a.y, b.y, s); s = fmaf(a.z, b.z, s); s = fmaf(a.w, b.w, s);
        amax = fmaxf(amax, fmaxf(fmaxf(fabsf(a.x), fabsf(a.y)), fmaxf(fabsf(a.z), fabsf(a.w))));
      }
    } else {
      for (int c = lane; c < C; c += 32) {
        const float a = dy[(size_t)m * dy_pitch + c];
        s = fmaf(a, o[(size_t)m * o_pitch + c], s);
        amax = fmaxf(amax, fabsf(a));
      }
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) s += __shfl_xor_sync(0xffffffffu, s, off);
    if (lane == 0) out[m] = s;
  }
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) amax = fmaxf(amax, __shfl_xor_sync(0xffffffffu, amax, off));
  if (lane == 0 && amax > 0.f && !(amax > 3.0e38f)) atomicMax(slot, __float_as_uint(amax));
}
// slot[0] = bits of max|dy|  ->  slot[1] = sc = 2^-(e+1) with max|dy| in [2^e, 2^(e+1)) (so |dy * sc| < 1), slot[2] = 1 / sc; both exact powers of two
__global__ void pam_scale_kernel(float* slot) {
  const unsigned bits = reinterpret_cast<unsigned*>(slot)[0];
  int e = (int)(bits >> 23) - 127;                  // 0 (all-zero gradient) gives e = -127: clamp below
  e = e < -100 ? -1 : (e > 100 ? 100 : e);
  slot[1] = exp2f((float)(-(e + 1)));
  slot[2] = exp2f((float)(e + 1));
}
// fp32 [rows][pitch] -> fp16 [rows][CPAD], zero padded, multiplied by the power-of-two scale *sc (sc == nullptr: 1).
// via_bf16: the values are rounded to bf16 first (and are then exact in fp16): the V operand of the backward must be the SAME rounded V the
// forward kernel multiplied P with, otherwise rowdot = sum dy*o (from the forward's o) and dP = dy V^T disagree by V's rounding and the
// cancellation in dS = P*(dP - rowdot) amplifies it (float64 emulation: dq 2.2e-3 with the consistent V, 6.0e-3 with fp16-rounded V).
// columns C, C+1: 1 (aug == nullptr, the V side) or -aug[row]*sc split into fp16 hi + lo (the dy side: -rowdot), so that dy V^T comes out as (dP - rowdot)*sc
__global__ void __launch_bounds__(256) pack_rows_f16_kernel(const float* __restrict__ src, int pitch, int C, long long rows, __half* __restrict__ dst,
                                                            const float* __restrict__ aug, const float* __restrict__ sc, int via_bf16) {
  const long long total = rows * (CPAD / 8);
  const float scale = sc ? __ldg(sc) : 1.f;
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
    const long long r = idx / (CPAD / 8); const int c0 = (int)(idx % (CPAD / 8)) * 8;
    __align__(16) __half h[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      const int c = c0 + e;
      float v = c < C ? __ldg(src + (size_t)r * pitch + c) * scale : 0.f;
      if (via_bf16) v = __bfloat162float(__float2bfloat16_rn(v));
      if (c == C || c == C + 1) {
        if (!aug) v = 1.f;
        else {
          const float a = -__ldg(aug + r) * scale;
          const float hi = __half2float(f2h_sat(a));
          v = c == C ? hi : a - hi;
        }
      }
      h[e] = f2h_sat(v);
    }
    *reinterpret_cast<uint4*>(dst + (size_t)r * CPAD + c0) = *reinterpret_cast<const uint4*>(h);
  }
}
// fp32 [B][N][d] (pitch) -> fp16 [B][DPAD][N], zero rows for c >= d
__global__ void __launch_bounds__(256) pack_t_f16_kernel(const float* __restrict__ v, int pitch, int d, int N, __half* __restrict__ vt) {
  __shared__ float tile[32][33];
  const int b = blockIdx.z, n0 = blockIdx.x * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  for (int i = ty; i < 32; i += 8) {
    const int n = n0 + i;
    tile[i][tx] = (n < N && tx < d) ? v[((size_t)b * N + n) * pitch + tx] : 0.f;
  }
  __syncthreads();
  for (int i = ty; i < 32; i += 8) {
    const int n = n0 + tx;
    if (n < N) vt[((size_t)b * DPAD + i) * N + n] = f2h_sat(tile[tx][i]);
  }
}
static int make_map_t(CUtensorMap* m, CUtensorMapDataType dt, const void* ptr, uint64_t rows, uint64_t cols, uint32_t box_rows, uint32_t box_cols, CUtensorMapSwizzle sw) {
  EncodeTiledFn enc = get_encode();
  if (!enc) { set_error("pam_tc: cuTensorMapEncodeTiled unavailable"); return GDN_ECUDA; }
  cuuint64_t gdim[2] = {cols, rows};
  cuuint64_t gstride[1] = {cols * 2};
  cuuint32_t box[2] = {box_cols, box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(m, dt, 2, const_cast<void*>(ptr), gdim, gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, sw, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { set_error("pam_tc: cuTensorMapEncodeTiled failed (%d)", (int)r); return GDN_ECUDA; }
  return GDN_OK;
}
}  // namespace bwd
}  // namespace pamtc
}  // namespace gdn

using namespace gdn::pamtc::bwd;

extern "C" int gdn_pam_tc_bwd_init(void) {
  GDN_CHECK_CUDA(cudaFuncSetAttribute(pam_flash_bwd_kernel<0, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, BwdSmem<0>::SMEM));
  GDN_CHECK_CUDA(cudaFuncSetAttribute(pam_flash_bwd_kernel<1, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, BwdSmem<0>::SMEM));
  GDN_CHECK_CUDA(cudaFuncSetAttribute(pam_flash_bwd_kernel<0, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, BwdSmem<1>::SMEM));
  GDN_CHECK_CUDA(cudaFuncSetAttribute(pam_flash_bwd_kernel<1, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, BwdSmem<1>::SMEM));
  return GDN_OK;
}

extern "C" size_t gdn_pam_tc_bwd_ws_bytes(const gdn_pam_bwd_args* a) {
  const size_t rows = (size_t)a->B * a->N;
  const size_t qk_row = DPAD * 2 * (a->precision == GDN_PREC_FP16X3 ? 2 : 1);
  return 256 + 2 * align256(rows * qk_row) + 2 * align256(rows * CPAD * 2) + 2 * align256((size_t)a->B * DPAD * a->N * 2);
}

// computes rowdot (= sum_c dy*o per row, an output: dgamma = sum rowdot) and the gradient scale itself
extern "C" int gdn_pam_tc_bwd(const gdn_pam_bwd_args* a, gdn_stream_t s) {
  GDN_CHECK_ARG(a->precision == GDN_PREC_FP16 || a->precision == GDN_PREC_FP16X3);
  const int split = a->precision == GDN_PREC_FP16X3;
  GDN_CHECK_ARG(a->N % TO == 0 && a->d + 2 <= DPAD && a->C + 2 <= CPAD && a->C % 4 == 0);    // two spare operand columns carry -lse / -rowdot
  GDN_CHECK_ARG(((uintptr_t)a->dv & 15) == 0 && ((uintptr_t)a->lse & 15) == 0 && ((uintptr_t)a->rowdot & 15) == 0);
  if (!a->ws || a->ws_bytes < gdn_pam_tc_bwd_ws_bytes(a)) { set_error("gdn_pam_bwd(fp16): workspace too small"); return GDN_EWORKSPACE; }
  const size_t rows = (size_t)a->B * a->N;
  const size_t qk_cols = (size_t)DPAD * (split ? 2 : 1);
  char* w = reinterpret_cast<char*>(a->ws);
  float* slot = reinterpret_cast<float*>(w); w += 256;          // [0] bits of max|dy|, [1] sc, [2] 1/sc
  __half* Qh = reinterpret_cast<__half*>(w); w += align256(rows * qk_cols * 2);
  __half* Kh = reinterpret_cast<__half*>(w); w += align256(rows * qk_cols * 2);
  __half* Vb = reinterpret_cast<__half*>(w); w += align256(rows * CPAD * 2);
  __half* DYb = reinterpret_cast<__half*>(w); w += align256(rows * CPAD * 2);
  __half* Qt = reinterpret_cast<__half*>(w); w += align256((size_t)a->B * DPAD * a->N * 2);
  __half* Kt = reinterpret_cast<__half*>(w);
  cudaStream_t st = as_stream(s);
  GDN_CHECK_CUDA(cudaMemsetAsync(slot, 0, 16, st));
  {
    const long long wb = cdiv((long long)rows, 8);
    rowdot_amax_kernel<<<(unsigned)(wb < 16 * kNumSMs ? wb : 16 * kNumSMs), 256, 0, st>>>(a->dy, a->dy_pitch, a->o, a->C, (long long)rows, a->C, a->rowdot,
                                                                                         reinterpret_cast<unsigned*>(slot));
    GDN_CHECK_LAUNCH();
    pam_scale_kernel<<<1, 1, 0, st>>>(slot);
    GDN_CHECK_LAUNCH();
  }
  const int pg = (int)(cdiv((long long)rows * DPAD, 256) < 16 * kNumSMs ? cdiv((long long)rows * DPAD, 256) : 16 * kNumSMs);
  pack_qk_kernel<<<pg, 256, 0, st>>>(a->q, a->qk_pitch, a->d, Qh, (long long)rows, a->lse, 1, split);       // columns d, d+1: -lse (hi, lo)
  GDN_CHECK_LAUNCH();
  pack_qk_kernel<<<pg, 256, 0, st>>>(a->k, a->qk_pitch, a->d, Kh, (long long)rows, nullptr, 1, split);      // columns d, d+1: 1
  GDN_CHECK_LAUNCH();
  const int rg = (int)(cdiv((long long)rows * (CPAD / 8), 256) < 16 * kNumSMs ? cdiv((long long)rows * (CPAD / 8), 256) : 16 * kNumSMs);
  pack_rows_f16_kernel<<<rg, 256, 0, st>>>(a->v, a->v_pitch, a->C, (long long)rows, Vb, nullptr, nullptr, 1);          // bf16-rounded v; columns C, C+1: 1
  GDN_CHECK_LAUNCH();
  pack_rows_f16_kernel<<<rg, 256, 0, st>>>(a->dy, a->dy_pitch, a->C, (long long)rows, DYb, a->rowdot, slot + 1, 0);    // dy*sc; columns C, C+1: -rowdot*sc (hi, lo)
  GDN_CHECK_LAUNCH();
  dim3 tg((unsigned)cdiv(a->N, 32), 1, (unsigned)a->B);
  pack_t_f16_kernel<<<tg, 256, 0, st>>>(a->q, a->qk_pitch, a->d, a->N, Qt);
  GDN_CHECK_LAUNCH();
  pack_t_f16_kernel<<<tg, 256, 0, st>>>(a->k, a->qk_pitch, a->d, a->N, Kt);
  GDN_CHECK_LAUNCH();
  CUtensorMap mq, mk, mv, mdy, mqt, mkt;
  int rc;
  if ((rc = make_map_t(&mq, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, Qh, rows, qk_cols, 64, DPAD, CU_TENSOR_MAP_SWIZZLE_64B)) != GDN_OK) return rc;
  if ((rc = make_map_t(&mk, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, Kh, rows, qk_cols, 64, DPAD, CU_TENSOR_MAP_SWIZZLE_64B)) != GDN_OK) return rc;
  if ((rc = make_map_t(&mv, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, Vb, rows, CPAD, 64, 64, CU_TENSOR_MAP_SWIZZLE_128B)) != GDN_OK) return rc;
  if ((rc = make_map_t(&mdy, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, DYb, rows, CPAD, 64, 64, CU_TENSOR_MAP_SWIZZLE_128B)) != GDN_OK) return rc;
  if ((rc = make_map_t(&mqt, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, Qt, (uint64_t)a->B * DPAD, (uint64_t)a->N, DPAD, 64, CU_TENSOR_MAP_SWIZZLE_128B)) != GDN_OK) return rc;
  if ((rc = make_map_t(&mkt, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, Kt, (uint64_t)a->B * DPAD, (uint64_t)a->N, DPAD, 64, CU_TENSOR_MAP_SWIZZLE_128B)) != GDN_OK) return rc;
  bwd::Params p;
  p.lse = a->lse; p.rowdot = a->rowdot; p.gamma = a->gamma; p.scale = slot; p.dq = a->dq; p.dk = a->dk; p.dv = a->dv;
  p.B = a->B; p.N = a->N; p.C = a->C; p.d = a->d;
  const int grid = a->B * (a->N / TO);
  if (split) {
    pam_flash_bwd_kernel<0, 1><<<grid, NTHREADS, BwdSmem<1>::SMEM, st>>>(mq, mk, mv, mdy, mqt, mkt, p);
    GDN_CHECK_LAUNCH();
    pam_flash_bwd_kernel<1, 1><<<grid, NTHREADS, BwdSmem<1>::SMEM, st>>>(mq, mk, mv, mdy, mqt, mkt, p);
  } else {
    pam_flash_bwd_kernel<0, 0><<<grid, NTHREADS, BwdSmem<0>::SMEM, st>>>(mq, mk, mv, mdy, mqt, mkt, p);
    GDN_CHECK_LAUNCH();
    pam_flash_bwd_kernel<1, 0><<<grid, NTHREADS, BwdSmem<0>::SMEM, st>>>(mq, mk, mv, mdy, mqt, mkt, p);
  }
  GDN_CHECK_LAUNCH();
  return GDN_OK;
}
