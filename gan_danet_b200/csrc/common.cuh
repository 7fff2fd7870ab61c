// Shared helpers for libgandanet_sm100.so (internal).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include "../../include/gandanet.h"

namespace gdn {

void set_error(const char* fmt, ...);

#define GDN_CHECK_ARG(cond)                                                          \
  do {                                                                               \
    if (!(cond)) {                                                                   \
      gdn::set_error("%s:%d: invalid argument: %s", __FILE__, __LINE__, #cond);      \
      return GDN_EINVAL;                                                             \
    }                                                                                \
  } while (0)

#define GDN_CHECK_CUDA(expr)                                                         \
  do {                                                                               \
    cudaError_t _e = (expr);                                                         \
    if (_e != cudaSuccess) {                                                         \
      gdn::set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #expr, cudaGetErrorString(_e)); \
      return GDN_ECUDA;                                                              \
    }                                                                                \
  } while (0)

#define GDN_CHECK_LAUNCH() GDN_CHECK_CUDA(cudaPeekAtLastError())

static inline cudaStream_t as_stream(gdn_stream_t s) { return reinterpret_cast<cudaStream_t>(s); }
static inline long long cdiv(long long a, long long b) { return (a + b - 1) / b; }

constexpr int kNumSMs = 148;

__device__ __forceinline__ float apply_act(float v, int act, float slope) {
  if (act == GDN_ACT_RELU) return v > 0.f ? v : 0.f;
  if (act == GDN_ACT_LRELU) return v > 0.f ? v : v * slope;
  return v;
}
// derivative factor given the pre-activation (or equivalently the output) sign
__device__ __forceinline__ float act_grad(float v, int act, float slope) {
  if (act == GDN_ACT_RELU) return v > 0.f ? 1.f : 0.f;
  if (act == GDN_ACT_LRELU) return v > 0.f ? 1.f : slope;
  return 1.f;
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// Block-wide sum (blockDim.x multiple of 32, <= 1024).  Result valid in thread 0.
template <typename T>
__device__ __forceinline__ T block_sum(T v, T* smem /* >= 32 */) {
  v = warp_sum(v);
  int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  if (lane == 0) smem[wid] = v;
  __syncthreads();
  int nw = (blockDim.x + 31) >> 5;
  v = (threadIdx.x < nw) ? smem[threadIdx.x] : T(0);
  if (wid == 0) v = warp_sum(v);
  __syncthreads();
  return v;
}

}  // namespace gdn
