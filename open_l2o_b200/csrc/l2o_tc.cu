// tcgen05 engine translation unit.
#include "cwlstm_ffma.cuh"   // load_vec / store_vec / preprocess helpers
#include "cwlstm_tc.cuh"
#include "cwlstm_tc_bwd.cuh"
#include "cwlstm_tc_bwd2.cuh"
#include <cstdlib>
#include <mutex>
#include <utility>
#include <vector>
#include "l2o_internal.h"

namespace l2o {
// forward / step: LSTM-20x2 with identity / LogAndSign preprocessing (cfg 0, 1) and RNNProp's fc(2->20)+ELU net (cfg 2);
// BPTT: cfg 0, 1 layer-pipelined in one kernel; cfg 2 as two single-chain passes (its K = 48 layer-1 operands do not fit next
// to layer 2's in shared memory / TMEM) with a caller-provided hand-over buffer
bool tc_supported(int cfg) { return cfg == 0 || cfg == 1 || cfg == 2; }
static bool tc_bwd_supported(int cfg) { return cfg == 0 || cfg == 1 || cfg == 2; }
bool tc_fwd_ok(const l2o_net* h, const l2o_unroll_args& a) {
  if (a.opt_kind == L2O_OPT_QUADRATIC_BATCH) return false;  // grouped optimizees exchange x: exact-fp32 engine only
  if (h->cfg == 2) return true;                             // fused Adam-feature mode (m, v) or given (m~, g~) rows
  return a.m == nullptr && a.feat_rec == nullptr;
}
// L2O_TC_AUTO=0: ENGINE_AUTO never picks the tcgen05 engine (A/B runs of whole tests against the exact-fp32 engine)
static bool tc_auto_env() {
  static const bool on = !(std::getenv("L2O_TC_AUTO") != nullptr && std::getenv("L2O_TC_AUTO")[0] == '0');
  return on;
}
bool tc_auto_default() { return tc_auto_env(); }
bool tc_bwd_auto_default() { return tc_auto_env(); }  // parity-green on the B200 (tests/test_tc_gpu.py)

// Weight-image buffers (2 x 97 KB each) are recycled through a process-wide free list and never cudaFree'd: a handle may
// be destroyed (Python GC) while ANOTHER program is capturing a CUDA graph, and cudaFree during a capture invalidates it.
namespace {
struct ImgPool {
  std::mutex mu;
  std::vector<std::pair<int, float*>> free_list;   // (device, pointer)
};
ImgPool& img_pool() {
  static ImgPool* p = new ImgPool();   // intentionally leaked: outlives every handle
  return *p;
}
}  // namespace

void tc_release_image(l2o_net* h) {
  if (!h->tc_img) return;
  ImgPool& P = img_pool();
  std::lock_guard<std::mutex> g(P.mu);
  P.free_list.emplace_back(h->tc_img_dev, h->tc_img);
  h->tc_img = nullptr;
}

static int ensure_image(l2o_net* h) {
  int dev = 0;
  L2O_CUDA_TRY(cudaGetDevice(&dev));
  if (h->tc_img == nullptr || h->tc_img_dev != dev) {
    tc_release_image(h);
    {
      ImgPool& P = img_pool();
      std::lock_guard<std::mutex> g(P.mu);
      for (size_t k = 0; k < P.free_list.size(); ++k)
        if (P.free_list[k].first == dev) {
          h->tc_img = P.free_list[k].second;
          P.free_list.erase(P.free_list.begin() + k);
          break;
        }
    }
    if (h->tc_img == nullptr) L2O_CUDA_TRY(cudaMalloc(&h->tc_img, 2 * tc::kImgAllBytes));   // fc nets: one BPTT image per pass
    h->tc_img_dev = dev;
    h->tc_img_mode = -1;
  }
  return L2O_OK;
}

bool tc_bwd_ok(const l2o_net* h, const l2o_bwd_args& a) {
  // meta-loss mode (lambda suffix sums of g_rec) or imitation mode with the forward pass's recorded deltas
  const bool mode_ok = a.labels ? (a.delta_seq != nullptr && a.n_total > 0) : a.g_rec != nullptr;
  if (h->cfg == 2 && a.scratch == nullptr) return false;   // fc nets: two passes with a caller-provided hand-over buffer
  if (h->rt.tanh_output && a.delta_seq == nullptr) return false;   // tanh' comes from the recorded deltas
  return tc_bwd_supported(h->cfg) && mode_ok;
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
    if (h->cfg == 2) rc = tc_launch_bwd2<Cfg<L2O_PRE_FC, 2, 20, 20, 20>>(h->rt, a, h->tc_img, st, sms);
  }
  if (rc == L2O_OK) count_launch(h->cfg == 2 ? 3 : 2);
  h->tc_img_mode = 1;
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
  const bool prep = !(s.reuse_weights && h->tc_img_mode == 0);   // forward image of this theta already in place
  if (h->cfg == 0) rc = tc_launch_fwd<Cfg<L2O_PRE_IDENTITY, 1, 1, 20, 20>>(h->rt, a, h->tc_img, st, sms, s.state_out, stage, ex, prep);
  if (h->cfg == 1) rc = tc_launch_fwd<Cfg<L2O_PRE_LOGSIGN, 1, 2, 20, 20>>(h->rt, a, h->tc_img, st, sms, s.state_out, stage, ex, prep);
  if (h->cfg == 2) rc = tc_launch_fwd<Cfg<L2O_PRE_FC, 2, 20, 20, 20>>(h->rt, a, h->tc_img, st, sms, s.state_out, stage, ex, prep);
  if (rc == L2O_OK) count_launch(prep ? 2 : 1);
  h->tc_img_mode = 0;
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
  h->tc_img_mode = 0;
  if (rc == L2O_E_CUDA) return set_cuda_error(cudaGetLastError(), "tc_unroll_fwd launch");
  return rc;
}
}  // namespace l2o
