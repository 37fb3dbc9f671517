// k_eval_small<NL, PV> instantiations: the one-launch logp + gradient evaluation of small batches
// (small.cuh).  The O(N) parameter maps are compiled once per unit, not per instantiation.  Built twice
// (launch_small.cu: Gaussian, launch_small_pv.cu: pseudo-Voigt) so the halves compile in parallel.
#include "launch.cuh"
#include "small.cuh"

#ifndef CHI_SMALL_PV
#define CHI_SMALL_PV 0
#endif

namespace chi {
#if CHI_SMALL_PV
int launch_eval_small_pv(int NL, const MapCfg& cfg, const SmallArgs& a, size_t dyn_smem, cudaStream_t st) {
#else
int launch_eval_small_pv(int NL, const MapCfg& cfg, const SmallArgs& a, size_t dyn_smem, cudaStream_t st);
static int launch_eval_small_g(int NL, const MapCfg& cfg, const SmallArgs& a, size_t dyn_smem, cudaStream_t st) {
#endif
  constexpr bool PV = CHI_SMALL_PV != 0;
  const unsigned grid = (unsigned)(a.B * a.n_chunks);
  switch (NL) {
    case 1:
      if (sm_allow_dyn(k_eval_small<1, PV>, dyn_smem)) return -2;
      k_eval_small<1, PV><<<grid, CHI_SM_THREADS, dyn_smem, st>>>(cfg, a);
      return 0;
    case 2:
      if (sm_allow_dyn(k_eval_small<2, PV>, dyn_smem)) return -2;
      k_eval_small<2, PV><<<grid, CHI_SM_THREADS, dyn_smem, st>>>(cfg, a);
      return 0;
  }
  return -1;
}
// the cluster form: the n_chunks blocks of a draw are one thread-block cluster (small.cuh, SM_CLUSTER)
namespace {
template <int NL, bool PV>
int run_cl(const MapCfg& cfg, const SmallArgs& a, size_t gb_bytes, cudaStream_t st) {
  const size_t x_off = (gb_bytes + 15) & ~(size_t)15;
  const size_t dyn = x_off + ((size_t)a.n_chunks * a.stride + (size_t)a.N * a.K * CHI_JAC_W + a.N) * sizeof(double);
  if (sm_allow_dyn(k_eval_small_cl<NL, PV>, dyn)) return -2;
  cudaLaunchConfig_t lc = {};
  lc.gridDim = dim3((unsigned)(a.B * a.n_chunks));
  lc.blockDim = dim3(CHI_SM_THREADS);
  lc.dynamicSmemBytes = dyn;
  lc.stream = st;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = (unsigned)a.n_chunks;
  at[0].val.clusterDim.y = 1;
  at[0].val.clusterDim.z = 1;
  lc.attrs = at;
  lc.numAttrs = 1;
  return cudaLaunchKernelEx(&lc, k_eval_small_cl<NL, PV>, cfg, a, (unsigned)x_off) == cudaSuccess ? 0 : -2;
}
}  // namespace
#if CHI_SMALL_PV
int launch_eval_small_cl_pv(int NL, const MapCfg& cfg, const SmallArgs& a, size_t gb_bytes, cudaStream_t st) {
#else
int launch_eval_small_cl_pv(int NL, const MapCfg& cfg, const SmallArgs& a, size_t gb_bytes, cudaStream_t st);
static int launch_eval_small_cl_g(int NL, const MapCfg& cfg, const SmallArgs& a, size_t gb_bytes, cudaStream_t st) {
#endif
  constexpr bool PV = CHI_SMALL_PV != 0;
  switch (NL) {
    case 1: return run_cl<1, PV>(cfg, a, gb_bytes, st);
    case 2: return run_cl<2, PV>(cfg, a, gb_bytes, st);
  }
  return -1;
}
#if !CHI_SMALL_PV
int launch_eval_small_cl(int NL, bool pv, const MapCfg& cfg, const SmallArgs& a, size_t gb_bytes, cudaStream_t st) {
  return pv ? launch_eval_small_cl_pv(NL, cfg, a, gb_bytes, st) : launch_eval_small_cl_g(NL, cfg, a, gb_bytes, st);
}
int launch_eval_small(int NL, bool pv, const MapCfg& cfg, const SmallArgs& a, size_t dyn_smem, cudaStream_t st) {
  return pv ? launch_eval_small_pv(NL, cfg, a, dyn_smem, st) : launch_eval_small_g(NL, cfg, a, dyn_smem, st);
}
// resident blocks per SM of the instantiation (its launch bounds)
int eval_small_blocks_per_sm(int NL) { return NL <= 2 ? 2 : 1; }
#endif
}  // namespace chi
