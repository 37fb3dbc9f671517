// Launchers of the two hot kernels, compiled in their own translation units
// (launch_em.cu / launch_abs.cu) so that the 8 x 2 template instantiations of each build in
// parallel.  Return 0, or -1 for an unsupported cloud count.
#pragma once
#include <cuda_runtime.h>

#include "moments.cuh"

namespace chi {
int launch_moments_em(int N, bool pv, const ChanArgs& a, cudaStream_t st, int family);
int launch_moments_abs(int N, bool pv, const ChanArgs& a, cudaStream_t st, int family);

inline unsigned warp_grid(const ChanArgs& a) {
  int64_t items = a.B * a.n_chunks;
  return (unsigned)((items + CHI_WARPS_PER_BLOCK - 1) / CHI_WARPS_PER_BLOCK);
}
}  // namespace chi
