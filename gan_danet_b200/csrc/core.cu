// Library plumbing (errors, init) and layout kernels of libgandanet_sm100.so.
// Layout note: the reference is NCHW fp32 throughout (models/generator.py, models/discriminator.py);
// this library works on NHWC fp32 slices, so the module boundary converts once on entry and once on exit.
#include <stdarg.h>
#include <string.h>
#include "common.cuh"

namespace gdn {
static thread_local char g_err[512] = "";
void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

// ---- NCHW -> NHWC through a 32x32 smem transpose of the (C, HW) plane of one sample
__global__ void nchw_to_nhwc_kernel(const float* __restrict__ src, float* __restrict__ dst, int dst_pitch, int C, int HW) {
  __shared__ float tile[32][33];
  const int b = blockIdx.z;
  const int p0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
  const float* s = src + (size_t)b * C * HW;
  float* d = dst + (size_t)b * HW * dst_pitch;
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    int c = c0 + i, p = p0 + threadIdx.x;
    if (c < C && p < HW) tile[i][threadIdx.x] = s[(size_t)c * HW + p];
  }
  __syncthreads();
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    int p = p0 + i, c = c0 + threadIdx.x;
    if (c < C && p < HW) d[(size_t)p * dst_pitch + c] = tile[threadIdx.x][i];
  }
}
__global__ void nhwc_to_nchw_kernel(const float* __restrict__ src, int src_pitch, float* __restrict__ dst, int C, int HW) {
  __shared__ float tile[32][33];
  const int b = blockIdx.z;
  const int p0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
  const float* s = src + (size_t)b * HW * src_pitch;
  float* d = dst + (size_t)b * C * HW;
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    int p = p0 + i, c = c0 + threadIdx.x;
    if (c < C && p < HW) tile[i][threadIdx.x] = s[(size_t)p * src_pitch + c];
  }
  __syncthreads();
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    int c = c0 + i, p = p0 + threadIdx.x;
    if (c < C && p < HW) d[(size_t)c * HW + p] = tile[threadIdx.x][i];
  }
}

// OIHW -> [O][kh][kw][I]  and -> [I][kh][kw][O]
__global__ void weight_perm_kernel(const float* __restrict__ w, float* __restrict__ out, int O, int I, int taps, int to_ihwo) {
  const long long total = (long long)O * I * taps;
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
    // idx enumerates the OUTPUT linearly (coalesced writes)
    int o, i, t;
    if (!to_ihwo) {
      i = (int)(idx % I); long long r = idx / I; t = (int)(r % taps); o = (int)(r / taps);
    } else {
      o = (int)(idx % O); long long r = idx / O; t = (int)(r % taps); i = (int)(r / taps);
    }
    out[idx] = __ldg(w + ((size_t)o * I + i) * taps + t);
  }
}

__global__ void fill_kernel(float* __restrict__ p, long long n, float v) {
  long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (; i < n; i += stride) p[i] = v;
}
}  // namespace gdn

using namespace gdn;

extern "C" int gdn_version(void) { return 100; }
extern "C" const char* gdn_last_error(void) { return g_err; }

extern "C" int gdn_pam_tc_init(void);   // pam_tc.cu: opts the tcgen05 kernels into large dynamic shared memory
extern "C" int gdn_conv_tc_init(void);  // conv_tc.cu
extern "C" int gdn_linear_tc_init(void);  // linear_tc.cu

// Per-device set-up (cudaFuncSetAttribute is per device).  The caller's current device is restored: initialising the library for cuda:1 must not
// move the process (and with it torch.cuda.current_device()) to cuda:1.
extern "C" int gdn_init(int device) {
  int prev = -1;
  GDN_CHECK_CUDA(cudaGetDevice(&prev));
  cudaDeviceProp prop;
  GDN_CHECK_CUDA(cudaGetDeviceProperties(&prop, device));
  if (prop.major != 10) {
    set_error("gdn_init: device %d is sm_%d%d, this library is sm_100a only", device, prop.major, prop.minor);
    return GDN_EARCH;
  }
  GDN_CHECK_CUDA(cudaSetDevice(device));
  int rc = gdn_pam_tc_init();
  if (rc == GDN_OK) rc = gdn_conv_tc_init();
  if (rc == GDN_OK) rc = gdn_linear_tc_init();
  if (prev >= 0 && prev != device) cudaSetDevice(prev);
  return rc;
}

extern "C" int gdn_nchw_to_nhwc(const float* src, float* dst, int dst_pitch, int dst_c0, int B, int C, int H, int W, gdn_stream_t s) {
  GDN_CHECK_ARG(src && dst && B > 0 && C > 0 && H > 0 && W > 0 && dst_pitch >= dst_c0 + C && B <= 65535);
  dim3 grid((unsigned)cdiv((long long)H * W, 32), (unsigned)cdiv(C, 32), (unsigned)B), block(32, 8);
  nchw_to_nhwc_kernel<<<grid, block, 0, as_stream(s)>>>(src, dst + dst_c0, dst_pitch, C, H * W);
  GDN_CHECK_LAUNCH();
  return GDN_OK;
}
extern "C" int gdn_nhwc_to_nchw(const float* src, int src_pitch, int src_c0, float* dst, int B, int C, int H, int W, gdn_stream_t s) {
  GDN_CHECK_ARG(src && dst && B > 0 && C > 0 && H > 0 && W > 0 && src_pitch >= src_c0 + C && B <= 65535);
  dim3 grid((unsigned)cdiv((long long)H * W, 32), (unsigned)cdiv(C, 32), (unsigned)B), block(32, 8);
  nhwc_to_nchw_kernel<<<grid, block, 0, as_stream(s)>>>(src + src_c0, src_pitch, dst, C, H * W);
  GDN_CHECK_LAUNCH();
  return GDN_OK;
}
static int weight_perm(const float* w, float* out, int O, int I, int kh, int kw, int to_ihwo, gdn_stream_t s) {
  GDN_CHECK_ARG(w && out && O > 0 && I > 0 && kh > 0 && kw > 0);
  long long total = (long long)O * I * kh * kw;
  int blocks = (int)(cdiv(total, 256) < 8 * kNumSMs ? cdiv(total, 256) : 8 * kNumSMs);
  weight_perm_kernel<<<blocks, 256, 0, as_stream(s)>>>(w, out, O, I, kh * kw, to_ihwo);
  GDN_CHECK_LAUNCH();
  return GDN_OK;
}
extern "C" int gdn_weight_oihw_to_ohwi(const float* w, float* out, int O, int I, int kh, int kw, gdn_stream_t s) {
  return weight_perm(w, out, O, I, kh, kw, 0, s);
}
extern "C" int gdn_weight_oihw_to_ihwo(const float* w, float* out, int O, int I, int kh, int kw, gdn_stream_t s) {
  return weight_perm(w, out, O, I, kh, kw, 1, s);
}
extern "C" int gdn_fill(float* p, long long n, float value, gdn_stream_t s) {
  GDN_CHECK_ARG(p && n >= 0);
  if (n == 0) return GDN_OK;
  int blocks = (int)(cdiv(n, 1024) < 8 * kNumSMs ? cdiv(n, 1024) : 8 * kNumSMs);
  fill_kernel<<<blocks, 256, 0, as_stream(s)>>>(p, n, value);
  GDN_CHECK_LAUNCH();
  return GDN_OK;
}
