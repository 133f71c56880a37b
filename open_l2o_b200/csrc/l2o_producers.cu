// Fused gradient producers for the synthetic optimizee families: the optimizee step either side
// of the hot path.  The reference evaluates f and df/dx with ~15 TF ops per unroll step (DM/problems.py:103-175 +
// tf.gradients at DM/meta.py:322-329); here ONE launch per step produces both, so the step-at-a-time regime is
// producer kernel -> l2o_step, all inside one captured CUDA graph per unroll.
//
// Lasso (DM/problems.py:103-135 `lasso`, :137-175 `lasso_fixed`):
//   f = mean_b( 0.5 * ||A_b x_b - y_b||^2 + l * ||x_b||_1 ),   df/dx_b = ( A_b^T (A_b x_b - y_b) + l * sign(x_b) ) / B
// One CTA per batch row b.  Pass 1: the residual r = A_b x_b - y_b (a warp per matrix row, lanes striding over the
// columns: coalesced 128-byte reads, shuffle reduction).  Pass 2: g_j = sum_i A_ij r_i (a thread per column, the
// residual broadcast from shared memory, A read row by row: coalesced).  A_b (500 KB at m=250, n=500) is streamed from
// L2 twice per step; the whole batch (64 MB at B=128) stays L2-resident across the unroll.
// Random-scaling trick (DM/meta_dm_train.py:336-338,384-385): with `scale` the loss is evaluated at x (.) scale and the
// chain rule multiplies the gradient by scale.
#include <cuda_runtime.h>

#include "l2o_internal.h"

namespace {

constexpr int kLassoThreads = 512;

__global__ void __launch_bounds__(kLassoThreads) lasso_grad_kernel(l2o_lasso_args a) {
  extern __shared__ __align__(16) float sm[];
  const int m = a.m, n = a.n;
  float* sx = sm;                 // [n]  x (.) scale
  float* sr = sm + ((n + 3) & ~3);  // [m]  residual
  __shared__ double red[kLassoThreads / 32];
  const int b = blockIdx.x;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const float* __restrict__ A = a.A + (size_t)b * m * n;
  const float* __restrict__ x = a.x + (size_t)b * n;
  const float* __restrict__ sc = a.scale ? a.scale + (size_t)b * n : nullptr;
  double l1 = 0.0;
  for (int j = tid; j < n; j += kLassoThreads) {
    const float xv = sc ? x[j] * sc[j] : x[j];
    sx[j] = xv;
    l1 += (double)fabsf(xv);
  }
  __syncthreads();
  // ---- pass 1: r_i = sum_j A_ij x_j - y_i -------------------------------------------------------------------------
  double sq = 0.0;
  for (int i = warp; i < m; i += kLassoThreads / 32) {
    const float* __restrict__ row = A + (size_t)i * n;
    float acc0 = 0.f, acc1 = 0.f;
    int j = lane;
    for (; j + 32 < n; j += 64) {
      acc0 = fmaf(row[j], sx[j], acc0);
      acc1 = fmaf(row[j + 32], sx[j + 32], acc1);
    }
    if (j < n) acc0 = fmaf(row[j], sx[j], acc0);
    float acc = acc0 + acc1;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if (lane == 0) {
      const float r = acc - a.y[(size_t)b * m + i];
      sr[i] = r;
      sq += (double)r * (double)r;
    }
  }
  __syncthreads();
  // ---- pass 2: g_j = (sum_i A_ij r_i + l sign(x_j)) / B  [* scale_j] ----------------------------------------------------
  const float inv_b = 1.0f / (float)a.batch;
  for (int j = tid; j < n; j += kLassoThreads) {
    float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
    int i = 0;
    for (; i + 4 <= m; i += 4) {
      a0 = fmaf(A[(size_t)i * n + j], sr[i], a0);
      a1 = fmaf(A[(size_t)(i + 1) * n + j], sr[i + 1], a1);
      a2 = fmaf(A[(size_t)(i + 2) * n + j], sr[i + 2], a2);
      a3 = fmaf(A[(size_t)(i + 3) * n + j], sr[i + 3], a3);
    }
    for (; i < m; ++i) a0 = fmaf(A[(size_t)i * n + j], sr[i], a0);
    const float xv = sx[j];
    const float sgn = xv > 0.f ? 1.f : (xv < 0.f ? -1.f : 0.f);   // tf.abs'(0) = sign(0) = 0
    float g = ((a0 + a1) + (a2 + a3) + a.l1 * sgn) * inv_b;
    if (sc) g *= sc[j];
    a.g[(size_t)b * n + j] = g;
  }
  // ---- f += (0.5 sum r^2 + l sum |x|) / B -----------------------------------------------------------------------------------
  if (a.f) {
    double part = 0.5 * sq + (double)a.l1 * l1;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) part += __shfl_xor_sync(0xffffffffu, part, o);
    if (lane == 0) red[warp] = part;
    __syncthreads();
    if (tid == 0) {
      double tot = 0.0;
      for (int w = 0; w < kLassoThreads / 32; ++w) tot += red[w];
      atomicAdd(a.f, tot / (double)a.batch);
    }
  }
}

}  // namespace

extern "C" int l2o_lasso_grad(const l2o_lasso_args* a, void* stream) {
  if (!a || a->batch < 0 || a->m <= 0 || a->n <= 0 || !a->A || !a->y || !a->x || !a->g) return L2O_E_INVALID;
  if (a->batch == 0) return L2O_OK;
  const size_t smem = (size_t)(((a->n + 3) & ~3) + a->m) * sizeof(float);
  if (smem > 200 * 1024) return L2O_E_UNSUPPORTED;
  L2O_CUDA_TRY(cudaFuncSetAttribute(lasso_grad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  lasso_grad_kernel<<<a->batch, kLassoThreads, smem, (cudaStream_t)stream>>>(*a);
  l2o::count_launch();
  L2O_CUDA_TRY(cudaGetLastError());
  return L2O_OK;
}
