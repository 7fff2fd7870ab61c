// Micro-benchmark: duration of back-to-back tcgen05.mma (kind::f16, M = 128, K = 16) instructions on one SM of sm_100a for
// different N, A-operand sources (shared memory / tensor memory) and B layouts (K-major / MN-major, SWIZZLE_128B).
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I../../gan_danet_b200/csrc -o umma_time umma_time.cu
#include <cstdio>
#include "tc_common.cuh"
namespace gdn { void set_error(const char*, ...) {} }
using namespace gdn::tc;
__device__ __forceinline__ uint64_t desc_mn(uint32_t addr, uint32_t lbo, uint32_t sbo) {
  return (uint64_t)((addr >> 4) & 0x3FFF) | ((uint64_t)((lbo >> 4) & 0x3FFF) << 16) | ((uint64_t)((sbo >> 4) & 0x3FFF) << 32) | ((uint64_t)1 << 46) | ((uint64_t)LAYOUT_SW128 << 61);
}
// mode bit0: A from TMEM; bit1: B MN-major
template <int N, int mode>
__global__ void __launch_bounds__(128) k(long long* out, int reps) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  __shared__ uint32_t tslot;
  __shared__ __align__(8) unsigned long long barmem;
  const uint32_t bar = smem_u32(&barmem);
  if (threadIdx.x == 0) { mbar_init(bar, 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
  if (threadIdx.x < 32) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tslot)), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before(); __syncthreads(); tc_fence_after();
  const uint32_t tmem = tslot;
  if (threadIdx.x < 32) {
    constexpr uint32_t idesc = idesc_f16(128, N) | ((mode & 2) ? (1u << 16) : 0u);
    uint64_t ad[8], bd[8];
#pragma unroll
    for (int ks = 0; ks < 8; ++ks) {
      ad[ks] = smem_desc(base + (ks >> 2) * 16384 + (ks & 3) * 32, 1024, LAYOUT_SW128);
      bd[ks] = (mode & 2) ? desc_mn(base + 32768 + (ks & 3) * 2048 + (ks >> 2) * 24576, 8192, 1024) : smem_desc(base + 32768 + (ks >> 2) * 32768 + (ks & 3) * 32, 1024, LAYOUT_SW128);
    }
    long long t0 = clock64();
    for (int r = 0; r < reps; ++r) {
#pragma unroll
      for (int ks = 0; ks < 8; ++ks) {
        if (elect_one()) {
        if (mode & 1) asm volatile("tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, 1;" ::"r"(tmem + 256), "r"(tmem + ks * 8), "l"(bd[ks]), "r"(idesc) : "memory");
        else asm volatile("tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, 1;" ::"r"(tmem + 256), "l"(ad[ks]), "l"(bd[ks]), "r"(idesc) : "memory");
        }
      }
    }
    if (elect_one()) tc_commit(bar);
    __syncwarp();
    mbar_wait(bar, 0);
    long long t1 = clock64();
    if (threadIdx.x == 0) out[blockIdx.x] = t1 - t0;
  }
  tc_fence_before(); __syncthreads();
  if (threadIdx.x < 32) { tc_fence_after(); asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u) : "memory"); }
}
int main() {
  long long* out; cudaMalloc(&out, 148 * 8);
  const int smem = 200 * 1024;
  const int reps = 64;
#define RUN(N, mode) { cudaFuncSetAttribute(k<N, mode>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem); k<N, mode><<<148, 128, smem>>>(out, reps); \
        cudaError_t e = cudaDeviceSynchronize(); long long h = 0; cudaMemcpy(&h, out, 8, cudaMemcpyDeviceToHost); \
        printf("N %3d A:%s B:%s  %.1f cycles per MMA (M128 x N x K16)  [%s]\n", N, (mode & 1) ? "tmem" : "smem", (mode & 2) ? "MN" : "K ", (double)h / (reps * 8), cudaGetErrorString(e)); }
  RUN(16, 0) RUN(16, 1) RUN(32, 0) RUN(32, 1) RUN(64, 0) RUN(64, 1) RUN(64, 3) RUN(128, 0) RUN(128, 1) RUN(192, 0) RUN(192, 1) RUN(192, 3) RUN(256, 0) RUN(256, 1)
  return 0;
}
