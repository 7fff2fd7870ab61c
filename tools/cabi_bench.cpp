// Python-free harness over the C ABI (SURVEY 8b "who calls it": the C benchmark harness, no torch anywhere): plain C++ host code
// that links libgandanet_sm100.so, owns its buffers through the CUDA runtime, calls gdn_pam_fwd / gdn_pam_bwd at the north-star shape
// (B = 32, N = 8192, C = 184, d = 23 by default) and
//   * checks a sample of output rows against a float64 host evaluation of generator.py:115-122 on the same inputs
//     (y = gamma * softmax(q k^T) v + x; the rows' softmax runs over all N keys),
//   * times the launches with CUDA events on the launching stream (W warm-up + K timed calls) and prints the algorithmic
//     TFLOP/s: 2*B*N^2*(d+C) forward, 4*B*N^2*(d+C) backward (operand packing inside the timed call).
// Build (gan_danet_b200/build.py does it): g++ tools/cabi_bench.cpp -Iinclude -I/usr/local/cuda/include -Lgan_danet_b200 -lgandanet_sm100
//        -L/usr/local/cuda/lib64 -lcudart -Wl,-rpath,'$ORIGIN' -o gan_danet_b200/cabi_bench
// Run:   gan_danet_b200/cabi_bench [B N C d]
#include <cuda_runtime.h>

#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "gandanet.h"

#define CK(x)                                                                         \
  do {                                                                                \
    cudaError_t e_ = (x);                                                             \
    if (e_ != cudaSuccess) { fprintf(stderr, "%s: %s\n", #x, cudaGetErrorString(e_)); return 2; } \
  } while (0)
#define GDN(x)                                                                        \
  do {                                                                                \
    int rc_ = (x);                                                                    \
    if (rc_ != GDN_OK) { fprintf(stderr, "%s failed (%d): %s\n", #x, rc_, gdn_last_error()); return 3; } \
  } while (0)

static uint64_t rng_state = 0x9E3779B97F4A7C15ull;
static float gauss() {   // Box-Muller on a splitmix64 stream: deterministic inputs without any library
  auto next = [] { uint64_t z = (rng_state += 0x9E3779B97F4A7C15ull); z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull; z = (z ^ (z >> 27)) * 0x94D049BB133111EBull; return z ^ (z >> 31); };
  const double u1 = ((next() >> 11) + 1.0) / 9007199254740993.0, u2 = (next() >> 11) / 9007199254740992.0;
  return (float)(std::sqrt(-2.0 * std::log(u1)) * std::cos(6.283185307179586 * u2));
}

int main(int argc, char** argv) {
  const int B = argc > 4 ? atoi(argv[1]) : 32, N = argc > 4 ? atoi(argv[2]) : 8192, C = argc > 4 ? atoi(argv[3]) : 184, d = argc > 4 ? atoi(argv[4]) : 23;
  const int warm = 3, iters = 10;
  GDN(gdn_init(0));
  const size_t rows = (size_t)B * N;
  std::vector<float> hq(rows * d), hk(rows * d), hv(rows * C), hx(rows * C), hdy(rows * C);
  for (auto& v : hq) v = 0.6f * gauss();
  for (auto& v : hk) v = 0.6f * gauss();
  for (auto& v : hv) v = gauss();
  for (auto& v : hx) v = gauss();
  for (auto& v : hdy) v = gauss();
  const float hgamma = 0.5f;
  float *q, *k, *v, *x, *y, *o, *lse, *gamma, *dy, *dq, *dk, *dv, *rowdot;
  CK(cudaMalloc(&q, rows * d * 4)); CK(cudaMalloc(&k, rows * d * 4)); CK(cudaMalloc(&v, rows * C * 4)); CK(cudaMalloc(&x, rows * C * 4));
  CK(cudaMalloc(&y, rows * C * 4)); CK(cudaMalloc(&o, rows * C * 4)); CK(cudaMalloc(&lse, rows * 4)); CK(cudaMalloc(&gamma, 4));
  CK(cudaMalloc(&dy, rows * C * 4)); CK(cudaMalloc(&dq, rows * d * 4)); CK(cudaMalloc(&dk, rows * d * 4)); CK(cudaMalloc(&dv, rows * C * 4));
  CK(cudaMalloc(&rowdot, rows * 4));
  CK(cudaMemcpy(q, hq.data(), rows * d * 4, cudaMemcpyHostToDevice)); CK(cudaMemcpy(k, hk.data(), rows * d * 4, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(v, hv.data(), rows * C * 4, cudaMemcpyHostToDevice)); CK(cudaMemcpy(x, hx.data(), rows * C * 4, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dy, hdy.data(), rows * C * 4, cudaMemcpyHostToDevice)); CK(cudaMemcpy(gamma, &hgamma, 4, cudaMemcpyHostToDevice));

  gdn_pam_fwd_args f = {};
  f.q = q; f.k = k; f.qk_pitch = d; f.d = d; f.v = v; f.v_pitch = C; f.x = x; f.x_pitch = C; f.gamma = gamma;
  f.o = o; f.y = y; f.y_pitch = C; f.lse = lse; f.B = B; f.N = N; f.C = C; f.precision = GDN_PREC_FP16; f.chunk = 0;
  gdn_pam_bwd_args g = {};
  g.q = q; g.k = k; g.qk_pitch = d; g.d = d; g.v = v; g.v_pitch = C; g.o = o; g.lse = lse; g.gamma = gamma; g.dy = dy; g.dy_pitch = C;
  g.dq = dq; g.dk = dk; g.dv = dv; g.rowdot = rowdot; g.B = B; g.N = N; g.C = C; g.precision = GDN_PREC_FP16; g.chunk = 0;
  size_t wsf = gdn_pam_fwd_ws_bytes(&f), wsb = gdn_pam_bwd_ws_bytes(&g);
  void* ws;
  CK(cudaMalloc(&ws, wsf > wsb ? wsf : wsb));
  f.ws = ws; f.ws_bytes = wsf; g.ws = ws; g.ws_bytes = wsb;
  cudaStream_t st;
  CK(cudaStreamCreate(&st));
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));

  auto time_ms = [&](auto&& call, float* ms) -> int {
    for (int i = 0; i < warm; ++i) GDN(call());
    CK(cudaStreamSynchronize(st));
    CK(cudaEventRecord(e0, st));
    for (int i = 0; i < iters; ++i) GDN(call());
    CK(cudaEventRecord(e1, st));
    CK(cudaStreamSynchronize(st));
    CK(cudaEventElapsedTime(ms, e0, e1));
    *ms /= iters;
    return 0;
  };
  float ms_f = 0, ms_b = 0;
  if (int rc = time_ms([&] { return gdn_pam_fwd(&f, st); }, &ms_f)) return rc;
  if (int rc = time_ms([&] { return gdn_pam_bwd(&g, st); }, &ms_b)) return rc;

  // ---- parity on sampled rows: float64 host evaluation of generator.py:115-122
  const int nsample = 24;
  std::vector<float> hy(rows * C);
  CK(cudaMemcpy(hy.data(), y, rows * C * 4, cudaMemcpyDeviceToHost));
  double num = 0, den = 0, anum = 0, aden = 0;
  std::vector<double> s(N), acc(C);
  for (int t = 0; t < nsample; ++t) {
    const int b = (t * 7) % B;
    const size_t i = (size_t)b * N + ((size_t)t * 2654435761u) % N;
    double m = -1e300;
    for (int j = 0; j < N; ++j) {
      double dot = 0;
      for (int c = 0; c < d; ++c) dot += (double)hq[i * d + c] * hk[((size_t)b * N + j) * d + c];
      s[j] = dot; if (dot > m) m = dot;
    }
    double z = 0;
    for (int c = 0; c < C; ++c) acc[c] = 0;
    for (int j = 0; j < N; ++j) {
      const double p = std::exp(s[j] - m);
      z += p;
      const float* vr = &hv[((size_t)b * N + j) * C];
      for (int c = 0; c < C; ++c) acc[c] += p * vr[c];
    }
    for (int c = 0; c < C; ++c) {
      const double att = acc[c] / z, want = hgamma * att + hx[i * C + c], got = hy[i * C + c];
      num += (got - want) * (got - want); den += want * want;
      const double gatt = (got - hx[i * C + c]) / hgamma;
      anum += (gatt - att) * (gatt - att); aden += att * att;
    }
  }
  const double err_y = std::sqrt(num / den), err_att = std::sqrt(anum / aden);
  const double flop_f = 2.0 * B * (double)N * N * (d + C), flop_b = 2.0 * flop_f;
  printf("{\"harness\": \"C ABI, no Python\", \"B\": %d, \"N\": %d, \"C\": %d, \"d\": %d, \"pam_fwd_ms\": %.4f, \"pam_fwd_tflops\": %.1f, \"pam_bwd_ms\": %.4f, "
         "\"pam_bwd_tflops\": %.1f, \"rows_checked\": %d, \"rel_err_y\": %.3e, \"rel_err_attention_term\": %.3e, \"abi_version\": %d}\n",
         B, N, C, d, ms_f, flop_f / ms_f / 1e9, ms_b, flop_b / ms_b / 1e9, nsample, err_y, err_att, gdn_version());
  return (err_y < 2e-3 && err_att < 6e-3) ? 0 : 1;
}
