// The two-launch form of the emission kernel for large batches (moments.cuh, MODE): EA_TAME over every draw
// without the range selects of exp(-tau), EA_FIXUP over the draws the cloud map marked.  Simple-baseline form
// only.  Built twice (launch_em_tame.cu: Gaussian, launch_em_tame_pv.cu: pseudo-Voigt) so the halves compile in
// parallel.
#include "launch.cuh"

#ifndef CHI_EM_TAME_PV
#define CHI_EM_TAME_PV 0
#endif

namespace chi {
namespace {
constexpr bool PV = CHI_EM_TAME_PV != 0;

int tame(int N, const ChanArgs& a, cudaStream_t st) {
  const unsigned grid = warp_grid(a);
  const bool staged = a.sl && a.staged;
  switch (N) {
#define CASE(n)                                                                                  \
  case n:                                                                                        \
    if (staged) k_moments_em<n, PV, true, true, EA_TAME><<<grid, CHI_BLOCK, 0, st>>>(a);          \
    else k_moments_em<n, PV, true, false, EA_TAME><<<grid, CHI_BLOCK, 0, st>>>(a);                \
    return 0;
    CASE(1) CASE(2) CASE(3) CASE(4) CASE(5) CASE(6) CASE(7) CASE(8)
#undef CASE
  }
  return -1;
}

int fixup(int N, const ChanArgs& a, int n_sm, cudaStream_t st) {
  switch (N) {  // one wave: as many blocks as stay resident
#define CASE(n)                                                                                                       \
  case n:                                                                                                             \
    k_moments_em<n, PV, true, false, EA_FIXUP><<<(unsigned)(n_sm * em_min_blocks(n, PV)), CHI_BLOCK, 0, st>>>(a);     \
    return 0;
    CASE(1) CASE(2) CASE(3) CASE(4) CASE(5) CASE(6) CASE(7) CASE(8)
#undef CASE
  }
  return -1;
}
}  // namespace

#if CHI_EM_TAME_PV
int launch_moments_em_tame_pv(int N, const ChanArgs& a, cudaStream_t st) { return tame(N, a, st); }
int launch_moments_em_fixup_pv(int N, const ChanArgs& a, int n_sm, cudaStream_t st) { return fixup(N, a, n_sm, st); }
#else
int launch_moments_em_tame_pv(int N, const ChanArgs& a, cudaStream_t st);
int launch_moments_em_fixup_pv(int N, const ChanArgs& a, int n_sm, cudaStream_t st);
bool moments_em_two_pass(int N, const ChanArgs& a) { return N <= 8 && a.n_base <= 1; }
int launch_moments_em_tame(int N, bool pv, const ChanArgs& a, cudaStream_t st) {
  return pv ? launch_moments_em_tame_pv(N, a, st) : tame(N, a, st);
}
int launch_moments_em_fixup(int N, bool pv, const ChanArgs& a, int n_sm, cudaStream_t st) {
  return pv ? launch_moments_em_fixup_pv(N, a, n_sm, st) : fixup(N, a, n_sm, st);
}
#endif
}  // namespace chi
