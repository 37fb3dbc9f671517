// STAGED (per-sightline spectra, bulk-async channel pipeline) instantiations of the fused kernel,
// simple-baseline form only; see moments_ea.cuh.
#include "launch.cuh"
#include "moments_ea.cuh"

namespace chi {
template <int N, bool PV>
static void run(const FusedArgs& a, cudaStream_t st) {
  const int64_t items = a.B * a.n_chunks;
  const unsigned grid = (unsigned)((items + CHI_WARPS_PER_BLOCK - 1) / CHI_WARPS_PER_BLOCK);
  k_moments_ea<N, PV, true, true><<<grid, CHI_BLOCK, 0, st>>>(a);
}

int launch_moments_ea_staged(int N, bool pv, const FusedArgs& a, cudaStream_t st) {
  switch (N * 2 + (pv ? 1 : 0)) {
#define CASE(n)                           \
  case (n)*2: run<n, false>(a, st); return 0; \
  case (n)*2 + 1: run<n, true>(a, st); return 0;
    CASE(1) CASE(2) CASE(3) CASE(4) CASE(5) CASE(6) CASE(7) CASE(8)
#undef CASE
  }
  return -1;
}

int launch_moments_ea_tame_staged(int N, const FusedArgs& a, cudaStream_t st) {
  const int64_t items = a.B * a.n_chunks;
  const unsigned grid = (unsigned)((items + CHI_WARPS_PER_BLOCK - 1) / CHI_WARPS_PER_BLOCK);
  switch (N) {
#define CASE(n) \
  case n: k_moments_ea<n, false, true, true, true, EA_TAME><<<grid, CHI_BLOCK, 0, st>>>(a); return 0;
    CASE(1) CASE(2) CASE(3) CASE(4) CASE(5) CASE(6) CASE(7) CASE(8)
#undef CASE
  }
  return -1;
}
}  // namespace chi
