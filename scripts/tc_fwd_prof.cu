// Timeline probe for the tcgen05 forward kernel (cwlstm_tc.cuh): clock64 stamps of CTA 0 for tile 0 half 0 / half 1, tile 1
// half 0 and the issuer, over a fused Rastrigin unroll (training layout: checkpoints written).  Prints per-phase durations
// of a few steady-state steps.
#define L2O_TC_FPROF 1
#define L2O_TC_FDBG 1
#include <unistd.h>
#include <cstdio>
#include <vector>
#include "cwlstm_ffma.cuh"
#include "cwlstm_tc.cuh"
using namespace l2o;
int main() {
  using C = Cfg<L2O_PRE_IDENTITY, 1, 1, 20, 20>;
  const int64_t n = 148 * 256 * 2; const int T = 12;
  float *theta, *state, *ckpt, *g_rec, *img, *x, *oa, *ob; double* fx;
  cudaMalloc(&theta, C::P * 4); cudaMalloc(&state, n * C::SF * 4); cudaMalloc(&g_rec, (T + 1) * n * 4);
  cudaMalloc(&ckpt, (size_t)(T + 1) * n * C::SF * 4); cudaMalloc(&img, 2 * tc::kImgAllBytes); cudaMalloc(&x, n * 4);
  cudaMalloc(&oa, n * 4); cudaMalloc(&ob, n * 4); cudaMalloc(&fx, (T + 1) * 8);
  std::vector<float> h(C::P); for (int i = 0; i < C::P; ++i) h[i] = 0.05f * ((i * 2654435761u % 1000) / 500.f - 1.f);
  cudaMemcpy(theta, h.data(), C::P * 4, cudaMemcpyHostToDevice);
  std::vector<float> hv(n); for (int64_t i = 0; i < n; ++i) hv[i] = ((i * 40503u % 2000) / 1000.f - 1.f);
  cudaMemcpy(x, hv.data(), n * 4, cudaMemcpyHostToDevice); cudaMemcpy(oa, hv.data(), n * 4, cudaMemcpyHostToDevice);
  cudaMemcpy(ob, hv.data(), n * 4, cudaMemcpyHostToDevice);
  cudaMemset(state, 0, n * C::SF * 4); cudaMemset(fx, 0, (T + 1) * 8);
  l2o_unroll_args a{}; a.n = n; a.T = T; a.theta = theta; a.opt_kind = L2O_OPT_RASTRIGIN_SEP; a.opt_a = oa; a.opt_b = ob;
  a.opt_alpha = 10.f; a.opt_fscale = 1.f / n; a.x = x; a.state = state; a.ckpt = ckpt; a.g_rec = g_rec; a.fx = fx; a.step0 = 1;
  NetRt rt{0.1f, 0.f, 1.f, 0};
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  int* dbg_h = nullptr;   // host-mapped progress words: read while a launch hangs
  cudaHostAlloc(&dbg_h, 64 * sizeof(int), cudaHostAllocMapped);
  for (int k = 0; k < 64; ++k) dbg_h[k] = -1;
  int* dbg_d = nullptr;
  cudaHostGetDevicePointer(&dbg_d, dbg_h, 0);
  cudaMemcpyToSymbol(tc::g_fdbg, &dbg_d, sizeof(dbg_d));
  for (int rep = 0; rep < 3; ++rep) {
    int zero[4] = {0, 0, 0, 0};
    cudaMemcpyToSymbol(tc::g_fprof_n, zero, sizeof(zero));
    cudaEventRecord(e0);
    int rc = tc_launch_fwd<C>(rt, a, img, 0, 148);
    cudaEventRecord(e1);
    for (int w = 0; w < 50 && cudaEventQuery(e1) == cudaErrorNotReady; ++w) usleep(100000);
    if (cudaEventQuery(e1) == cudaErrorNotReady) {
      printf("HANG in launch %d; progress words (kpair<<16 | t<<4 | stage):\n", rep);
      const char* nm[8] = {"t0h0 step", "t0h0 staged", "", "", "t0h1 step", "t0h1 staged", "", ""};
      for (int k = 0; k < 12; ++k) if (k % 4 < 2) printf("  [%d] %s = 0x%x\n", k, k < 8 ? nm[k] : (k == 8 ? "t1h0 step" : "t1h0 staged"), dbg_h[k]);
      printf("  store warp: picked t0 0x%x t1 0x%x  done t0 0x%x t1 0x%x\n", dbg_h[16], dbg_h[17], dbg_h[18], dbg_h[19]);
      printf("  issuer: t0 ev %d t1 ev %d   kernel entered %d\n", dbg_h[20], dbg_h[21], dbg_h[30]);
      fflush(stdout);
      _exit(3);
    }
    cudaError_t e = cudaDeviceSynchronize();
    if (rc || e != cudaSuccess) { printf("rc=%d err=%s\n", rc, cudaGetErrorString(e)); return 1; }
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    printf("launch %d: %.3f ms for %lld coords x T=%d (2 pairs per SM) -> %.0f cycles per pair-step at 1.9 GHz\n", rep, ms,
           (long long)n, T, ms * 1e-3 * 1.9e9 / (2.0 * T));
  }
  std::vector<long long> p(4 * 4096);
  cudaMemcpyFromSymbol(p.data(), tc::g_fprof, sizeof(long long) * 4 * 4096);
  auto tm = [&](int role, int idx) { return p[role * 4096 + idx] >> 3; };
  const char* nm[7] = {"start", "A arrived", "D1 ready", "epi1 + B arrived", "D2 ready", "epi2", "pair barrier"};
  const char* rn[3] = {"tile 0 half 0 (12 units)", "tile 0 half 1 (8 units + scalars)", "tile 1 half 0"};
  for (int role = 0; role < 3; ++role) {
    printf("== %s: cycles since previous stamp, steps 4..7 of the first pair\n", rn[role]);
    for (int st = 4; st < 8; ++st) {
      printf(" step %d:", st);
      for (int e = 0; e < 7; ++e) printf(" %s=%lld", nm[e], tm(role, st * 7 + e) - tm(role, st * 7 + e - 1));
      printf(" | step total %lld\n", tm(role, st * 7 + 6) - tm(role, (st - 1) * 7 + 6));
    }
  }
  printf("== absolute stamps (cycles since tile-0 half-0 step-4 start)\n");
  const long long t0 = tm(0, 4 * 7);
  for (int st = 4; st < 7; ++st)
    for (int role = 0; role < 3; ++role) {
      printf(" role %d step %d:", role, st); for (int e = 0; e < 7; ++e) printf(" %lld", tm(role, st * 7 + e) - t0); printf("\n");
    }
  printf("== issuer events (tile.A/B @ cycles since the same origin)\n  ");
  for (int k = 0; k < 4096; ++k) {
    const long long dt = tm(3, k) - t0;
    if (dt > -3000 && dt < 16000) printf(" t%lld%c@%lld", (p[3 * 4096 + k] >> 1) & 1, (p[3 * 4096 + k] & 1) ? 'B' : 'A', dt);
  }
  printf("\n");
  return 0;
}
