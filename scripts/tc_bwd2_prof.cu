// Timeline probe for the layer-pipelined tcgen05 BPTT kernel (cwlstm_tc_bwd2.cuh): clock64 stamps of CTA 0 for the
// layer-2 worker, the layer-1 worker and the issuer; prints per-phase durations for a few steady-state steps.
#define L2O_TC_PROF2 1
#include <cstdio>
#include <vector>
#include "cwlstm_ffma.cuh"
#include "cwlstm_tc_bwd2.cuh"
using namespace l2o;
int main() {
  using C = Cfg<L2O_PRE_IDENTITY, 1, 1, 20, 20>;
  const int64_t n = 148 * 128 * 2; const int T = 10;
  float *theta, *in_seq, *ckpt, *g_rec, *img; double* dth;
  cudaMalloc(&theta, C::P * 4); cudaMalloc(&in_seq, (T + 1) * n * 4); cudaMalloc(&g_rec, (T + 1) * n * 4);
  cudaMalloc(&ckpt, (size_t)(T + 1) * n * C::SF * 4); cudaMalloc(&dth, C::P * 8); cudaMalloc(&img, tc::kImgAllBytes);
  std::vector<float> h(C::P); for (int i = 0; i < C::P; ++i) h[i] = 0.05f * ((i * 2654435761u % 1000) / 500.f - 1.f);
  cudaMemcpy(theta, h.data(), C::P * 4, cudaMemcpyHostToDevice);
  cudaMemset(in_seq, 0, (T + 1) * n * 4); cudaMemset(g_rec, 0, (T + 1) * n * 4);
  cudaMemset(ckpt, 0, (size_t)(T + 1) * n * C::SF * 4); cudaMemset(dth, 0, C::P * 8);
  l2o_bwd_args a{}; a.n = n; a.T = T; a.theta = theta; a.in_seq = in_seq; a.ckpt = ckpt; a.g_rec = g_rec; a.dtheta = dth;
  NetRt rt{1.f, 0.f, 1.f, 0};
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  for (int rep = 0; rep < 3; ++rep) {
    cudaEventRecord(e0);
    int rc = tc_launch_bwd2<C>(rt, a, img, 0, 148);
    cudaEventRecord(e1);
    cudaError_t e = cudaDeviceSynchronize();
    if (rc || e != cudaSuccess) { printf("rc=%d err=%s\n", rc, cudaGetErrorString(e)); return 1; }
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    printf("launch %d: %.3f ms for %lld coords x T=%d (2 tiles per SM) -> %.0f cycles per tile-step at 1.9 GHz\n", rep, ms,
           (long long)n, T, ms * 1e-3 * 1.9e9 / (2.0 * T));
  }
  std::vector<long long> p(3 * 4096);
  cudaMemcpyFromSymbol(p.data(), tcb2::g_prof2, sizeof(long long) * 3 * 4096);
  auto tm = [&](int role, int idx) { return p[role * 4096 + idx] >> 3; };
  const char* e2[6] = {"start", "x_done+dh2c", "A2 arrived", "w_done", "X2 staged + z_done", "phase (dz staged), arrived"};
  printf("== layer-2 worker (warp 0 lane 0): cycles since previous stamp, steps 3..6\n");
  for (int st = 3; st < 7; ++st) {
    printf(" step %d:", st);
    for (int e = 0; e < 6; ++e) printf(" %s=%lld", e2[e], tm(0, st * 6 + e) - tm(0, st * 6 + e - 1));
    printf(" | step total %lld\n", tm(0, st * 6 + 5) - tm(0, (st - 1) * 6 + 5));
  }
  const char* e1n[7] = {"start", "x_done1+dh1c", "A1 arrived", "w_done", "X1 staged + x_done2+read", "z_done", "phase, arrived"};
  printf("== layer-1 worker (warp 8 lane 0): cycles since previous stamp, steps 3..6\n");
  for (int st = 3; st < 7; ++st) {
    printf(" step %d:", st);
    for (int e = 0; e < 7; ++e) printf(" %s=%lld", e1n[e], tm(1, st * 7 + e) - tm(1, st * 7 + e - 1));
    printf(" | step total %lld\n", tm(1, st * 7 + 6) - tm(1, (st - 1) * 7 + 6));
  }
  const char* ev[4] = {"Z2", "Z1", "dX2+dW2", "dX1+dW1"};
  printf("== issuer events, absolute (cycles since L2 step-3 start)\n  ");
  for (int k = 4; k < 40; ++k) { const long long dt = tm(2, k) - tm(0, 3 * 6); if (dt > -12000 && dt < 40000) printf(" %s@%lld", ev[p[2 * 4096 + k] & 3], dt); }
  printf("\n== interleaving: absolute stamps (cycles since L2 step-3 start)\n");
  const long long t0 = tm(0, 3 * 6);
  for (int st = 3; st < 6; ++st) {
    printf(" L2 step %d:", st); for (int e = 0; e < 6; ++e) printf(" %lld", tm(0, st * 6 + e) - t0); printf("\n");
    printf(" L1 step %d:", st); for (int e = 0; e < 7; ++e) printf(" %lld", tm(1, st * 7 + e) - t0); printf("\n");
  }
  return 0;
}
