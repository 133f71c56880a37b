// tcgen05 engine translation unit.
#include "cwlstm_ffma.cuh"   // load_vec / store_vec / preprocess helpers
#include "cwlstm_tc.cuh"
#include "cwlstm_tc_bwd.cuh"
#include "cwlstm_tc_bwd2.cuh"
#include <cstdlib>
#include "l2o_internal.h"

namespace l2o {
// forward / step: LSTM-20x2 with identity / LogAndSign preprocessing (cfg 0, 1) and RNNProp's fc(2->20)+ELU net (cfg 2);
// BPTT: cfg 0, 1 (RNNProp's K = 48 weight images do not fit next to the dW staging)
bool tc_supported(int cfg) { return cfg == 0 || cfg == 1 || cfg == 2; }
static bool tc_bwd_supported(int cfg) { return cfg == 0 || cfg == 1; }
bool tc_fwd_ok(const l2o_net* h, const l2o_unroll_args& a) {
  if (a.opt_kind == L2O_OPT_QUADRATIC_BATCH) return false;  // grouped optimizees exchange x: exact-fp32 engine only
  if (h->cfg == 2) return true;                             // fused Adam-feature mode (m, v) or given (m~, g~) rows
  return a.m == nullptr && a.feat_rec == nullptr;
}
bool tc_auto_default() { return true; }
bool tc_bwd_auto_default() { return true; }  // parity-green on the B200 (tests/test_tc_gpu.py)

static int ensure_image(l2o_net* h) {
  int dev = 0;
  L2O_CUDA_TRY(cudaGetDevice(&dev));
  if (h->tc_img == nullptr || h->tc_img_dev != dev) {
    if (h->tc_img) cudaFree(h->tc_img);
    h->tc_img = nullptr;
    L2O_CUDA_TRY(cudaMalloc(&h->tc_img, tc::kImgAllBytes));
    h->tc_img_dev = dev;
  }
  return L2O_OK;
}

bool tc_bwd_ok(const l2o_net* h, const l2o_bwd_args& a) {
  // meta-loss mode (lambda suffix sums of g_rec) or imitation mode with the forward pass's recorded deltas
  const bool mode_ok = a.labels ? (a.delta_seq != nullptr && a.n_total > 0) : a.g_rec != nullptr;
  return tc_bwd_supported(h->cfg) && mode_ok && !h->rt.tanh_output;
}

int tc_unroll_bwd(l2o_net* h, const l2o_bwd_args& a, cudaStream_t st) {
  if (!tc_bwd_ok(h, a)) return L2O_E_UNSUPPORTED;
  int rc = ensure_image(h);
  if (rc) return rc;
  const int sms = device_sms();
  if (sms <= 0) return L2O_E_CUDA;
  rc = L2O_E_UNSUPPORTED;
  // L2O_BWD_V1=1 selects the first-generation (phase-serial) kernel for A/B measurements; the layer-pipelined kernel
  // (cwlstm_tc_bwd2.cuh) is the product path
  static const bool v1 = std::getenv("L2O_BWD_V1") != nullptr && std::getenv("L2O_BWD_V1")[0] == '1';
  if (v1) {
    if (h->cfg == 0) rc = tc_launch_bwd<Cfg<L2O_PRE_IDENTITY, 1, 1, 20, 20>>(h->rt, a, h->tc_img, st, sms);
    if (h->cfg == 1) rc = tc_launch_bwd<Cfg<L2O_PRE_LOGSIGN, 1, 2, 20, 20>>(h->rt, a, h->tc_img, st, sms);
  } else {
    if (h->cfg == 0) rc = tc_launch_bwd2<Cfg<L2O_PRE_IDENTITY, 1, 1, 20, 20>>(h->rt, a, h->tc_img, st, sms);
    if (h->cfg == 1) rc = tc_launch_bwd2<Cfg<L2O_PRE_LOGSIGN, 1, 2, 20, 20>>(h->rt, a, h->tc_img, st, sms);
  }
  if (rc == L2O_OK) count_launch(2);
  if (rc == L2O_E_CUDA) return set_cuda_error(cudaGetLastError(), "tc_unroll_bwd launch");
  return rc;
}

bool tc_step_ok(const l2o_net* h, const l2o_step_args& a) {
  if (h->cfg == 2) return a.m != nullptr;   // fused RNNProp features; precomputed (m~, g~) pairs stay on the FFMA engine
  return tc_supported(h->cfg) && a.m == nullptr && a.in1 == nullptr && a.feat_out == nullptr;
}

// One time step with the state in HBM (external-gradient regime) = the forward unroll with T = 1 and an
// out-of-place final-state write.
int tc_step(l2o_net* h, const l2o_step_args& s, cudaStream_t st) {
  if (!tc_step_ok(h, s)) return L2O_E_UNSUPPORTED;
  int rc = ensure_image(h);
  if (rc) return rc;
  const int sms = device_sms();
  if (sms <= 0) return L2O_E_CUDA;
  l2o_unroll_args a{};
  a.n = s.n;
  a.T = 1;
  a.theta = s.theta;
  a.in_seq = s.in0;
  a.opt_kind = L2O_OPT_NONE;
  a.x = s.x;
  a.state = const_cast<float*>(s.state_in);
  a.delta_seq = s.delta;
  a.m = s.m;
  a.v = s.v;
  a.beta1 = s.beta1;
  a.beta2 = s.beta2;
  a.step0 = 1;
  a.feat_rec = s.feat_out;
  const tc::FwdExtra ex{s.step_ptr, s.t_offset, s.step_ptr ? 0.f : s.p};
  rc = L2O_E_UNSUPPORTED;
  // L2O_STEP_STAGE=0 disables the TMA-staged state loads (A/B measurements)
  static const bool stage = !(std::getenv("L2O_STEP_STAGE") != nullptr && std::getenv("L2O_STEP_STAGE")[0] == '0');
  if (h->cfg == 0) rc = tc_launch_fwd<Cfg<L2O_PRE_IDENTITY, 1, 1, 20, 20>>(h->rt, a, h->tc_img, st, sms, s.state_out, stage);
  if (h->cfg == 1) rc = tc_launch_fwd<Cfg<L2O_PRE_LOGSIGN, 1, 2, 20, 20>>(h->rt, a, h->tc_img, st, sms, s.state_out, stage);
  if (h->cfg == 2) rc = tc_launch_fwd<Cfg<L2O_PRE_FC, 2, 20, 20, 20>>(h->rt, a, h->tc_img, st, sms, s.state_out, stage, ex);
  if (rc == L2O_OK) count_launch(2);
  if (rc == L2O_E_CUDA) return set_cuda_error(cudaGetLastError(), "tc_step launch");
  return rc;
}

int tc_unroll_fwd(l2o_net* h, const l2o_unroll_args& a, cudaStream_t st) {
  if (!tc_supported(h->cfg) || !tc_fwd_ok(h, a)) return L2O_E_UNSUPPORTED;
  {
    int rc0 = ensure_image(h);
    if (rc0) return rc0;
  }
  const int sms = device_sms();
  if (sms <= 0) return L2O_E_CUDA;
  int rc = L2O_E_UNSUPPORTED;
  if (h->cfg == 0) rc = tc_launch_fwd<Cfg<L2O_PRE_IDENTITY, 1, 1, 20, 20>>(h->rt, a, h->tc_img, st, sms);
  if (h->cfg == 1) rc = tc_launch_fwd<Cfg<L2O_PRE_LOGSIGN, 1, 2, 20, 20>>(h->rt, a, h->tc_img, st, sms);
  if (h->cfg == 2) rc = tc_launch_fwd<Cfg<L2O_PRE_FC, 2, 20, 20, 20>>(h->rt, a, h->tc_img, st, sms);
  if (rc == L2O_OK) count_launch(2);
  if (rc == L2O_E_CUDA) return set_cuda_error(cudaGetLastError(), "tc_unroll_fwd launch");
  return rc;
}
}  // namespace l2o
