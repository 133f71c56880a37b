// bf16 MN-major SWIZZLE_128B staging check for the layer-pipelined BPTT kernel (cwlstm_tc_bwd2.cuh): fills a Y buffer
// through tcb2::y16_off with small integers (exact in bf16), runs the 8 K-steps of D = Y^T.Y (M = 128 slots, N = 48,
// K = 128 coordinates, tcgen05.mma.kind::f16 SS-mode with the kernel's own descriptors) and compares every D[m][n]
// with the integer dot product.  Also times back-to-back SS bf16 MMAs (N = 48 / 32).
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "cwlstm_ffma.cuh"
#include "cwlstm_tc_bwd2.cuh"
using namespace l2o;
using namespace l2o::tc;
using namespace l2o::tcb2;

__host__ __device__ inline int yval(int slot, int c) { return ((slot * 7 + c * 3 + (slot * c) % 5) % 17) - 8; }

__global__ void __launch_bounds__(160, 1) probe(int n, int* bad, float* first, float* timing, int reps) {
  extern __shared__ __align__(1024) unsigned char raw_[];
  unsigned char* raw = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(raw_) + 1023) & ~uintptr_t(1023));
  unsigned char* yh = raw;                     // 32 KB
  uint64_t* bars = reinterpret_cast<uint64_t*>(raw + 32768);
  uint32_t* slot = reinterpret_cast<uint32_t*>(bars + 4);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int k = threadIdx.x; k < 32768 / 4; k += blockDim.x) reinterpret_cast<uint32_t*>(yh)[k] = 0u;
  __syncthreads();
  // every (slot, coord): one bf16 value at y16_off
  for (int e = threadIdx.x; e < 128 * 128; e += blockDim.x) {
    const int s = e >> 7, c = e & 127;
    const __nv_bfloat16 v = __float2bfloat16((float)yval(s, c));
    *reinterpret_cast<__nv_bfloat16*>(yh + y16_off(c, s)) = v;
  }
  if (warp == 4) {
    if (lane == 0) { mbar_init(&bars[0], 1); mbar_init(&bars[1], 1); fence_barrier_init(); }
    __syncwarp();
    tmem_alloc(slot, 512);
    tmem_relinquish();
  }
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tb = *slot;
  if (warp == 4) {
    const bool leader = elect_one();
    const uint64_t yd = tcb::make_desc(smem_u32(yh), kLBO16, kSBO16, 2);
    const uint32_t id = make_idesc_bf16(n, 1, 1);
    constexpr uint64_t kStep = (2 * kSBO16) >> 4;
    if (leader) {
      for (int kb = 0; kb < 8; ++kb) mma_bf16_ss(tb, yd + kb * kStep, yd + kb * kStep, id, kb > 0 ? 1u : 0u);
      tc_commit(&bars[0]);
    }
    __syncwarp();
    mbar_wait(&bars[0], 0);
    if (reps > 0) {
      const long long t0 = clock64();
      for (int r = 0; r < reps; r += 8) {
#pragma unroll
        for (int q = 0; q < 8; ++q)
          if (leader) mma_bf16_ss(tb + 256, yd + q * kStep, yd + q * kStep, id, 1u);
      }
      const long long t1 = clock64();
      if (leader) tc_commit(&bars[1]);
      __syncwarp();
      mbar_wait(&bars[1], 0);
      const long long t2 = clock64();
      if (leader) { timing[0] = (float)(t1 - t0) / reps; timing[1] = (float)(t2 - t0) / reps; }
    }
  } else {
    mbar_wait(&bars[0], 0);
    tc_fence_after();
    const int m = warp * 32 + lane;
    const uint32_t td = tb + ((uint32_t)(warp * 32) << 16);
    int nbad = 0;
    for (int c4 = 0; c4 < n / 4; ++c4) {
      float v[4];
      tcb::tmem_ld4(td + 4 * c4, v);
      for (int e = 0; e < 4; ++e) {
        const int nn = 4 * c4 + e;
        int want = 0;
        for (int c = 0; c < 128; ++c) want += yval(m, c) * yval(nn, c);
        if (v[e] != (float)want) {
          if (nbad == 0 && m < 128) { first[3 * m] = (float)nn; first[3 * m + 1] = v[e]; first[3 * m + 2] = (float)want; }
          ++nbad;
        }
      }
    }
    bad[m] = nbad;
    tc_fence_before();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (warp == 4) tmem_dealloc(tb, 512);
}

int main(int argc, char** argv) {
  const int n = argc > 1 ? atoi(argv[1]) : 48, reps = argc > 2 ? atoi(argv[2]) : 0;
  int* dbad; float *dfirst, *dtim;
  cudaMalloc(&dbad, 128 * 4); cudaMalloc(&dfirst, 128 * 3 * 4); cudaMalloc(&dtim, 8);
  cudaMemset(dfirst, 0, 128 * 3 * 4); cudaMemset(dtim, 0, 8);
  const size_t smem = 32768 + 256 + 1024;
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  probe<<<1, 160, smem>>>(n, dbad, dfirst, dtim, reps);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("CUDA error %s\n", cudaGetErrorString(e)); return 1; }
  std::vector<int> bad(128); std::vector<float> first(384); float tim[2];
  cudaMemcpy(bad.data(), dbad, 512, cudaMemcpyDeviceToHost);
  cudaMemcpy(first.data(), dfirst, 1536, cudaMemcpyDeviceToHost);
  cudaMemcpy(tim, dtim, 8, cudaMemcpyDeviceToHost);
  int tot = 0, rows = 0;
  for (int m = 0; m < 128; ++m) { tot += bad[m]; rows += bad[m] > 0; }
  printf("bf16 MN-major SW128 probe N=%d: %d mismatching outputs in %d rows -> %s\n", n, tot, rows, tot == 0 ? "LAYOUT OK" : "LAYOUT MISMATCH");
  for (int m = 0, shown = 0; m < 128 && shown < 8; ++m)
    if (bad[m]) { printf("  row %d: %d bad, first at n=%d got %.1f want %.1f\n", m, bad[m], (int)first[3 * m], first[3 * m + 1], first[3 * m + 2]); ++shown; }
  if (reps > 0) printf("THROUGHPUT bf16 SS M128 N%d K16, %d back-to-back: issue %.1f cycles/MMA, issue+complete %.1f cycles/MMA\n", n, reps, tim[0], tim[1]);
  return tot == 0 ? 0 : 2;
}
