// k_moments_abs<N, PV, SB> instantiations (N = 1..8, Gaussian / pseudo-Voigt).  Built twice
// (launch_abs.cu: CHI_SB=1, launch_abs_gb.cu: CHI_SB=0) so the two halves compile in parallel.
#include "launch.cuh"

#ifndef CHI_SB
#define CHI_SB 1
#endif

namespace chi {
template <int N, bool PV>
static void run(const ChanArgs& a, cudaStream_t st) {
  k_moments_abs<N, PV, CHI_SB != 0><<<warp_grid(a), CHI_BLOCK, 0, st>>>(a);
}

#if CHI_SB
static int launch_sb(int N, bool pv, const ChanArgs& a, cudaStream_t st) {
#else
int launch_moments_abs_gb(int N, bool pv, const ChanArgs& a, cudaStream_t st) {
#endif
  N = padded_clouds(N);
  switch (N * 2 + (pv ? 1 : 0)) {
#define CASE(n)                        \
  case (n)*2: run<n, false>(a, st); return 0; \
  case (n)*2 + 1: run<n, true>(a, st); return 0;
    CASE(1) CASE(2) CASE(3) CASE(4) CASE(5) CASE(6) CASE(7) CASE(8)
#if !CHI_SB
    CASE(12) CASE(16)  // N = 9..16 run padded on these (general-baseline half only)
#endif
#undef CASE
  }
  return -1;
}

#if CHI_SB
int launch_moments_abs_gb(int N, bool pv, const ChanArgs& a, cudaStream_t st);
int launch_moments_abs_staged(int N, bool pv, const ChanArgs& a, cudaStream_t st);
int launch_moments_abs(int N, bool pv, const ChanArgs& a, cudaStream_t st) {
  // per-draw sightlines (survey mode): no L1 reuse of the channel data -> staged pipeline
  if (a.sl && a.staged && a.n_base <= 1 && N <= 8) return launch_moments_abs_staged(N, pv, a, st);
  return (a.n_base <= 1 && N <= 8) ? launch_sb(N, pv, a, st) : launch_moments_abs_gb(N, pv, a, st);
}
#endif
}  // namespace chi
