// k_moments_ea<N, PV, SB> instantiations: fused emission+absorption for a shared spectral axis.
// Built twice (launch_ea.cu: CHI_EA_SB=1, launch_ea_gb.cu: CHI_EA_SB=0) so the two halves compile
// in parallel.
#include <cstdlib>

#include "launch.cuh"
#include "moments_ea.cuh"

#ifndef CHI_EA_SB
#define CHI_EA_SB 1
#endif

namespace chi {
template <int N, bool PV>
static void run(const FusedArgs& a, cudaStream_t st) {
  const int64_t items = a.B * a.n_chunks;
  const unsigned grid = (unsigned)((items + CHI_WARPS_PER_BLOCK - 1) / CHI_WARPS_PER_BLOCK);
#if CHI_EA_SB
  static const int novjp = [] { const char* e = getenv("CHI_EA_NOVJP"); return e ? atoi(e) : 1; }();
  if (!a.vjp && (novjp == 2 || (novjp == 1 && PV))) {  // logp + gradient: instantiation without the VJP plumbing
    k_moments_ea<N, PV, true, false, false><<<grid, CHI_BLOCK, 0, st>>>(a);
    return;
  }
#endif
  k_moments_ea<N, PV, CHI_EA_SB != 0><<<grid, CHI_BLOCK, 0, st>>>(a);
}

#if CHI_EA_SB
int launch_moments_ea_sb(int N, bool pv, const FusedArgs& a, cudaStream_t st) {
#else
int launch_moments_ea_gb(int N, bool pv, const FusedArgs& a, cudaStream_t st) {
#endif
  N = padded_clouds(N);
  switch (N * 2 + (pv ? 1 : 0)) {
#define CASE(n)                           \
  case (n)*2: run<n, false>(a, st); return 0; \
  case (n)*2 + 1: run<n, true>(a, st); return 0;
    CASE(1) CASE(2) CASE(3) CASE(4) CASE(5) CASE(6) CASE(7) CASE(8)
#if !CHI_EA_SB
    CASE(12) CASE(16)  // N = 9..16 run padded on these (general-baseline half only)
#endif
#undef CASE
  }
  return -1;
}

#if CHI_EA_SB
int launch_moments_ea_gb(int N, bool pv, const FusedArgs& a, cudaStream_t st);
int launch_moments_ea_staged(int N, bool pv, const FusedArgs& a, cudaStream_t st);
int launch_moments_ea(int N, bool pv, const FusedArgs& a, cudaStream_t st) {
  // per-draw sightlines (survey mode): no L1 reuse of the channel data -> staged pipeline
  if (a.sl && a.staged && a.nb_e <= 1 && a.nb_a <= 1 && N <= 8) return launch_moments_ea_staged(N, pv, a, st);
  return (a.nb_e <= 1 && a.nb_a <= 1 && N <= 8) ? launch_moments_ea_sb(N, pv, a, st) : launch_moments_ea_gb(N, pv, a, st);
}
#endif
}  // namespace chi
