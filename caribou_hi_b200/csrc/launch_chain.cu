// k_nuts_chain<NL, PV> instantiations (nuts_chain.cuh).  Built twice (launch_chain.cu: Gaussian,
// launch_chain_pv.cu: pseudo-Voigt) so the halves compile in parallel.
#include "nuts_chain.cuh"

#ifndef CHI_CHAIN_PV
#define CHI_CHAIN_PV 0
#endif

namespace chi {
namespace {
template <int NL, bool PV>
int run_chain(int threads, const MapCfg& cfg, const SmallArgs& a, const NutsState& s, const NutsAsync& A, int max_ticks,
              size_t gb_bytes, cudaStream_t st) {
  if (threads == CHI_SPEC_THREADS) {  // the state machine on a ninth warp, next to the evaluation
    if (sm_allow_dyn(k_nuts_chain_spec<NL, PV>, gb_bytes)) return -2;
    k_nuts_chain_spec<NL, PV><<<(unsigned)a.B, CHI_SPEC_THREADS, gb_bytes, st>>>(cfg, a, s, A, max_ticks);
  } else if (threads == 512) {
    if (sm_allow_dyn(k_nuts_chain<NL, PV, 512>, gb_bytes)) return -2;
    k_nuts_chain<NL, PV, 512><<<(unsigned)a.B, 512, gb_bytes, st>>>(cfg, a, s, A, max_ticks);
  } else {
    if (sm_allow_dyn(k_nuts_chain<NL, PV, CHI_SM_THREADS>, gb_bytes)) return -2;
    k_nuts_chain<NL, PV, CHI_SM_THREADS><<<(unsigned)a.B, CHI_SM_THREADS, gb_bytes, st>>>(cfg, a, s, A, max_ticks);
  }
  return 0;
}
}  // namespace

// the cluster form (k_nuts_chain_cl): a.n_chunks blocks per chain
namespace {
template <int NL, bool PV>
int run_chain_cl(const MapCfg& cfg, const SmallArgs& a, const NutsState& s, const NutsAsync& A, int max_ticks, size_t gb_bytes,
                 cudaStream_t st) {
  const size_t x_off = (gb_bytes + 15) & ~(size_t)15;
  const size_t dyn = x_off + ((size_t)a.n_chunks * a.stride + (size_t)a.N * a.K * CHI_JAC_W + a.N) * sizeof(double);
  if (sm_allow_dyn(k_nuts_chain_cl<NL, PV>, dyn)) return -2;
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
  return cudaLaunchKernelEx(&lc, k_nuts_chain_cl<NL, PV>, cfg, a, s, A, max_ticks, (unsigned)x_off) == cudaSuccess ? 0 : -2;
}
}  // namespace
#if CHI_CHAIN_PV
int launch_nuts_chain_cl_pv(int NL, const MapCfg& cfg, const SmallArgs& a, const NutsState& s, const NutsAsync& A, int max_ticks,
                            size_t gb_bytes, cudaStream_t st) {
#else
int launch_nuts_chain_cl_pv(int NL, const MapCfg& cfg, const SmallArgs& a, const NutsState& s, const NutsAsync& A, int max_ticks,
                            size_t gb_bytes, cudaStream_t st);
static int launch_nuts_chain_cl_g(int NL, const MapCfg& cfg, const SmallArgs& a, const NutsState& s, const NutsAsync& A,
                                  int max_ticks, size_t gb_bytes, cudaStream_t st) {
#endif
  switch (NL) {
    case 1: return run_chain_cl<1, CHI_CHAIN_PV != 0>(cfg, a, s, A, max_ticks, gb_bytes, st);
    case 2: return run_chain_cl<2, CHI_CHAIN_PV != 0>(cfg, a, s, A, max_ticks, gb_bytes, st);
  }
  return -1;
}
#if !CHI_CHAIN_PV
int launch_nuts_chain_cl(int NL, bool pv, const MapCfg& cfg, const SmallArgs& a, const NutsState& s, const NutsAsync& A,
                         int max_ticks, size_t gb_bytes, cudaStream_t st) {
  return pv ? launch_nuts_chain_cl_pv(NL, cfg, a, s, A, max_ticks, gb_bytes, st)
            : launch_nuts_chain_cl_g(NL, cfg, a, s, A, max_ticks, gb_bytes, st);
}
#endif

#if CHI_CHAIN_PV
int launch_nuts_chain_pv(int NL, int threads, const MapCfg& cfg, const SmallArgs& a, const NutsState& s, const NutsAsync& A, int max_ticks,
                         size_t gb_bytes, cudaStream_t st) {
#else
int launch_nuts_chain_pv(int NL, int threads, const MapCfg& cfg, const SmallArgs& a, const NutsState& s, const NutsAsync& A, int max_ticks,
                         size_t gb_bytes, cudaStream_t st);
static int launch_nuts_chain_g(int NL, int threads, const MapCfg& cfg, const SmallArgs& a, const NutsState& s, const NutsAsync& A,
                               int max_ticks, size_t gb_bytes, cudaStream_t st) {
#endif
  constexpr bool PV = CHI_CHAIN_PV != 0;
  switch (NL) {
    case 1: return run_chain<1, PV>(threads, cfg, a, s, A, max_ticks, gb_bytes, st);
    case 2: return run_chain<2, PV>(threads, cfg, a, s, A, max_ticks, gb_bytes, st);
  }
  return -1;
}
#if !CHI_CHAIN_PV
int launch_nuts_chain(int NL, int threads, bool pv, const MapCfg& cfg, const SmallArgs& a, const NutsState& s, const NutsAsync& A,
                      int max_ticks, size_t gb_bytes, cudaStream_t st) {
  return pv ? launch_nuts_chain_pv(NL, threads, cfg, a, s, A, max_ticks, gb_bytes, st)
            : launch_nuts_chain_g(NL, threads, cfg, a, s, A, max_ticks, gb_bytes, st);
}
#endif
}  // namespace chi
