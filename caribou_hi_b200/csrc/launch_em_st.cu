// STAGED (per-sightline spectra, bulk-async channel pipeline) instantiations of k_moments_em,
// simple-baseline form only; see moments.cuh.
#include "launch.cuh"

namespace chi {
template <int N, bool PV>
static void run(const ChanArgs& a, cudaStream_t st) {
  k_moments_em<N, PV, true, true><<<warp_grid(a), CHI_BLOCK, 0, st>>>(a);
}

int launch_moments_em_staged(int N, bool pv, const ChanArgs& a, cudaStream_t st) {
  switch (N * 2 + (pv ? 1 : 0)) {
#define CASE(n)                        \
  case (n)*2: run<n, false>(a, st); return 0; \
  case (n)*2 + 1: run<n, true>(a, st); return 0;
    CASE(1) CASE(2) CASE(3) CASE(4) CASE(5) CASE(6) CASE(7) CASE(8)
#undef CASE
  }
  return -1;
}
}  // namespace chi
