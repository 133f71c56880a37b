// Timeline probe for the tcgen05 BPTT kernel: runs it on synthetic data with L2O_TC_PROF stamps and prints per-phase
// durations (cycles) of CTA 0 for a few steps.  Debug tool; not part of the library.
#define L2O_TC_PROF 1
#include <cstdio>
#include <vector>
#include "cwlstm_ffma.cuh"
#include "cwlstm_tc_bwd.cuh"
using namespace l2o;
int main() {
  using C = Cfg<L2O_PRE_IDENTITY, 1, 1, 20, 20>;
  const int64_t n = 148 * 128 * 2; const int T = 8;
  float *theta, *in_seq, *ckpt, *g_rec, *img; double* dth;
  cudaMalloc(&theta, C::P * 4); cudaMalloc(&in_seq, (T + 1) * n * 4); cudaMalloc(&g_rec, (T + 1) * n * 4);
  cudaMalloc(&ckpt, (size_t)(T + 1) * n * C::SF * 4); cudaMalloc(&dth, C::P * 8); cudaMalloc(&img, tc::kImgAllBytes);
  std::vector<float> h(C::P); for (int i = 0; i < C::P; ++i) h[i] = 0.05f * ((i * 2654435761u % 1000) / 500.f - 1.f);
  cudaMemcpy(theta, h.data(), C::P * 4, cudaMemcpyHostToDevice);
  cudaMemset(in_seq, 0, (T + 1) * n * 4); cudaMemset(g_rec, 0, (T + 1) * n * 4);
  cudaMemset(ckpt, 0, (size_t)(T + 1) * n * C::SF * 4); cudaMemset(dth, 0, C::P * 8);
  l2o_bwd_args a{}; a.n = n; a.T = T; a.theta = theta; a.in_seq = in_seq; a.ckpt = ckpt; a.g_rec = g_rec; a.dtheta = dth;
  NetRt rt{1.f, 0.f, 1.f, 0};
  for (int rep = 0; rep < 2; ++rep) {
    int rc = tc_launch_bwd<C>(rt, a, img, 0, 148);
    cudaError_t e = cudaDeviceSynchronize();
    if (rc || e != cudaSuccess) { printf("rc=%d err=%s\n", rc, cudaGetErrorString(e)); return 1; }
  }
  std::vector<long long> p(3 * 2048);
  cudaMemcpyFromSymbol(p.data(), tcb::g_prof, sizeof(long long) * 3 * 2048);
  const char* en[9] = {"loads issued", "dX1(prev) ready+carry", "P0 arrived", "Z1Z2 ready", "w_done(dW1 prev)", "P2 arrived",
                       "dX2 ready", "w_done(dW2)", "P3 arrived"};
  for (int role = 0; role < 2; ++role) {
    printf("== epilogue half %d (warp q=0 lane 0): cycles since previous event, steps 2..4 of tile 0\n", role);
    for (int st = 2; st < 5; ++st) {
      printf(" step %d:", st);
      for (int e = 0; e < 9; ++e) {
        const int idx = st * 9 + e;
        printf(" %s=%lld", en[e], p[role * 2048 + idx] - p[role * 2048 + idx - 1]);
      }
      printf("  | step total %lld\n", p[role * 2048 + st * 9 + 8] - p[role * 2048 + (st - 1) * 9 + 8]);
    }
  }
  const char* in[8] = {"a_ready(P0)", "Z1Z2 issued", "a_ready(P2)", "dX2 issued", "dW2 issued", "a_ready(P3)", "dX1 issued", "dW1 issued"};
  printf("== issuer: cycles since previous event\n");
  for (int st = 2; st < 5; ++st) {
    printf(" step %d:", st);
    for (int e = 0; e < 8; ++e) { const int idx = st * 8 + e; printf(" %s=%lld", in[e], p[2 * 2048 + idx] - p[2 * 2048 + idx - 1]); }
    printf("\n");
  }
  return 0;
}
