// Internal (non-ABI) declarations shared by the translation units of libl2o_b200.so.
#pragma once
#include <cuda_runtime.h>

#include "cwlstm_common.cuh"
#include "l2o_b200.h"

// Net shapes compiled into this build.  (id, PRE, NIN, F, H1, H2)
#define L2O_FOR_EACH_CFG(X)               \
  X(0, L2O_PRE_IDENTITY, 1, 1, 20, 20)    \
  X(1, L2O_PRE_LOGSIGN, 1, 2, 20, 20)     \
  X(2, L2O_PRE_FC, 2, 20, 20, 20)         \
  X(3, L2O_PRE_IDENTITY, 1, 1, 0, 0)      \
  X(4, L2O_PRE_IDENTITY, 1, 1, 1, 0)      \
  X(5, L2O_PRE_IDENTITY, 1, 1, 1, 1)      \
  X(6, L2O_PRE_IDENTITY, 1, 1, 2, 3)

struct l2o_net {
  l2o_net_desc desc;
  int cfg;
  int engine;
  int64_t n_theta;
  int64_t state_floats;
  l2o::NetRt rt;
  float* tc_img;   // device-side weight image of the tcgen05 engine (owned; lazily allocated)
  int tc_img_dev;
  int tc_img_mode; // what the image currently holds: -1 nothing, 0 forward layout, 1 BPTT layout
};

namespace l2o {
int set_cuda_error(cudaError_t e, const char* where);  // records the message, returns L2O_E_CUDA
void count_launch(int n = 1);
int device_sms();  // SM count of the current device (cached), 0 on failure

int ffma_step(const l2o_net* h, const l2o_step_args& a, cudaStream_t st);
int ffma_unroll_fwd(const l2o_net* h, const l2o_unroll_args& a, cudaStream_t st);
int ffma_unroll_bwd(const l2o_net* h, const l2o_bwd_args& a, cudaStream_t st);

bool tc_supported(int cfg);
void tc_release_image(l2o_net* h);   // hand the weight image back to the process-wide pool (never cudaFree)
bool tc_fwd_ok(const l2o_net* h, const l2o_unroll_args& a);
int tc_unroll_fwd(l2o_net* h, const l2o_unroll_args& a, cudaStream_t st);
bool tc_step_ok(const l2o_net* h, const l2o_step_args& a);
int tc_step(l2o_net* h, const l2o_step_args& a, cudaStream_t st);
bool tc_auto_default();
bool tc_bwd_auto_default();
bool tc_bwd_ok(const l2o_net* h, const l2o_bwd_args& a);
int tc_unroll_bwd(l2o_net* h, const l2o_bwd_args& a, cudaStream_t st);  // does ENGINE_AUTO pick the tcgen05 engine when it can?
}  // namespace l2o

#define L2O_CUDA_TRY(expr)                                              \
  do {                                                                  \
    cudaError_t e__ = (expr);                                           \
    if (e__ != cudaSuccess) return l2o::set_cuda_error(e__, #expr);     \
  } while (0)
