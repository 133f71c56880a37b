// FFMA engine, BPTT kernel instantiations.
#include "l2o_ffma_launch.cuh"

namespace l2o {
template <class C>
static int do_unroll_bwd(const l2o_net* h, const l2o_bwd_args& a, cudaStream_t st) {
  auto k = unroll_bwd_kernel<C>;
  const size_t smem = BwdGeom<C>::BYTES;
  int grid = 1;
  int rc = ffma_launch_cfg(k, smem, a.n, grid);
  if (rc) return rc;
  k<<<grid, kTile, smem, st>>>(a, h->rt);
  count_launch();
  L2O_CUDA_TRY(cudaGetLastError());
  return L2O_OK;
}

int ffma_unroll_bwd(const l2o_net* h, const l2o_bwd_args& a, cudaStream_t st) {
#define X(id, PRE, NIN, F, H1, H2) \
  if (h->cfg == id) return do_unroll_bwd<Cfg<PRE, NIN, F, H1, H2>>(h, a, st);
  L2O_FOR_EACH_CFG(X)
#undef X
  return L2O_E_UNSUPPORTED;
}
}  // namespace l2o
