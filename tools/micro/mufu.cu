// Micro-benchmark: MUFU.EX2 throughput per SM sub-partition on sm_100a, alone and in the softmax instruction mix
// (FFMA + EX2 + FADD + half2 pack).  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o mufu mufu.cu && ./mufu
#include <cstdio>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
__device__ __forceinline__ float ex2(float x) { float y; asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
template <int MODE>
__global__ void k(float* out, long long* cyc, int iters, float seed) {
  float v[32];
#pragma unroll
  for (int i = 0; i < 32; ++i) v[i] = seed + 0.001f * (threadIdx.x + i);
  float acc0 = 0.f, acc1 = 0.f;
  unsigned pk = 0;
  __syncthreads();
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 32; i += 2) {
      if (MODE == 0) { v[i] = ex2(v[i]); v[i + 1] = ex2(v[i + 1]); }
      else {
        float p0 = ex2(fmaf(v[i], 1.4426950408889634f, -seed)), p1 = ex2(fmaf(v[i + 1], 1.4426950408889634f, -seed));
        acc0 += p0; acc1 += p1;
        __half2 h = __floats2half2_rn(p0, p1);
        pk ^= *reinterpret_cast<unsigned*>(&h);
        if (MODE == 2) { v[i] += 1e-7f; v[i + 1] += 1e-7f; }
      }
    }
  }
  long long t1 = clock64();
  float s = acc0 + acc1 + __uint_as_float(pk);
#pragma unroll
  for (int i = 0; i < 32; ++i) s += v[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
int main() {
  float* out; long long* cyc;
  cudaMalloc(&out, 148 * 1024 * 4); cudaMalloc(&cyc, 148 * 8);
  const int iters = 2000;
  for (int mode = 0; mode < 2; ++mode)
    for (int threads : {128, 256, 512}) {
      if (mode == 0) k<0><<<148, threads>>>(out, cyc, iters, 0.5f); else k<1><<<148, threads>>>(out, cyc, iters, 0.5f);
      cudaDeviceSynchronize();
      long long h[148]; cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
      double c = (double)h[0];
      double warps_per_smsp = threads / 128.0;
      printf("mode %d threads %d: %.0f cycles; %.2f cycles per EX2 warp-instruction per SMSP (%.2f warps/SMSP)\n", mode, threads, c, c / (iters * 32.0 * warps_per_smsp), warps_per_smsp);
    }
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
