// The two-launch form of the fused kernel for large Gaussian batches (moments_ea.cuh, MODE): EA_TAME over
// every draw without the +inf branch of 2^y, EA_FIXUP over the draws the cloud map marked.  Simple-baseline
// form only; the staged EA_TAME instantiations are in launch_ea_st.cu.
#include "launch.cuh"
#include "moments_ea.cuh"

namespace chi {
int launch_moments_ea_tame_staged(int N, const FusedArgs& a, cudaStream_t st);

bool moments_ea_two_pass(int N, bool pv, const FusedArgs& a) {
  return !pv && N <= 8 && a.nb_e <= 1 && a.nb_a <= 1;
}

int launch_moments_ea_tame(int N, const FusedArgs& a, cudaStream_t st) {
  if (a.sl && a.staged) return launch_moments_ea_tame_staged(N, a, st);
  const int64_t items = a.B * a.n_chunks;
  const unsigned grid = (unsigned)((items + CHI_WARPS_PER_BLOCK - 1) / CHI_WARPS_PER_BLOCK);
  switch (N) {
#define CASE(n) \
  case n: k_moments_ea<n, false, true, false, true, EA_TAME><<<grid, CHI_BLOCK, 0, st>>>(a); return 0;
    CASE(1) CASE(2) CASE(3) CASE(4) CASE(5) CASE(6) CASE(7) CASE(8)
#undef CASE
  }
  return -1;
}

int launch_moments_ea_fixup(int N, const FusedArgs& a, int n_sm, cudaStream_t st) {
  switch (N) {  // one wave: as many blocks as stay resident
#define CASE(n)                                                                                                  \
  case n:                                                                                                        \
    k_moments_ea<n, false, true, false, true, EA_FIXUP><<<(unsigned)(n_sm * ea_min_blocks(n, false)), CHI_BLOCK, 0, st>>>(a); \
    return 0;
    CASE(1) CASE(2) CASE(3) CASE(4) CASE(5) CASE(6) CASE(7) CASE(8)
#undef CASE
  }
  return -1;
}
}  // namespace chi
