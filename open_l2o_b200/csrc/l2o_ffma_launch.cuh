// Launch-geometry helper of the FFMA engine translation units.
#pragma once
#include "cwlstm_ffma.cuh"
#include "l2o_internal.h"

namespace l2o {
template <class K>
int ffma_launch_cfg(K kernel, size_t smem, int64_t n, int& grid) {
  L2O_CUDA_TRY(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  int occ = 0;
  L2O_CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kernel, kTile, smem));
  if (occ < 1) return L2O_E_UNSUPPORTED;
  const int sms = device_sms();
  if (sms <= 0) return L2O_E_CUDA;
  const int64_t tiles = (n + kTile - 1) / kTile;
  const int64_t cap = (int64_t)sms * occ;
  grid = (int)(tiles < cap ? tiles : cap);
  if (grid < 1) grid = 1;
  return L2O_OK;
}
}  // namespace l2o
