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
  if (threads == 512) {
    if (sm_allow_dyn(k_nuts_chain<NL, PV, 512>, gb_bytes)) return -2;
    k_nuts_chain<NL, PV, 512><<<(unsigned)a.B, 512, gb_bytes, st>>>(cfg, a, s, A, max_ticks);
  } else {
    if (sm_allow_dyn(k_nuts_chain<NL, PV, CHI_SM_THREADS>, gb_bytes)) return -2;
    k_nuts_chain<NL, PV, CHI_SM_THREADS><<<(unsigned)a.B, CHI_SM_THREADS, gb_bytes, st>>>(cfg, a, s, A, max_ticks);
  }
  return 0;
}
}  // namespace

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
