// Launchers of the hot kernels, compiled in their own translation units (launch_*.cu) so that
// the 8 x 2 x 2 template instantiations of each build in parallel.  Return 0, or -1 for an unsupported cloud count.
#pragma once
#include <cuda_runtime.h>

#include "moments.cuh"

namespace chi {
// argument block of the spectra-only kernels
struct SpecArgs {
  int64_t B;
  int N;
  int S;
  int n_base;
  int pv;
  int emission;  // 1: radiative transfer, 0: absorption
  double bg;
  const double* __restrict__ vax;
  const double* __restrict__ basis;
  const double* __restrict__ offset;  // [S] or null
  const double* __restrict__ coeff;   // [B][n_base] or null
  const double* __restrict__ cc;
  double* __restrict__ mu;            // [B][S]
  // second dataset of the fused shared-axis kernel (launch_spectra_ea): absorption
  int n_base_a;
  const double* __restrict__ basis_a;
  const double* __restrict__ offset_a;
  const double* __restrict__ coeff_a;
  double* __restrict__ mu_a;
};
// argument block of the single-precision spectra kernels (launch_spectra_f32.cu)
struct Spec32Args {
  int64_t B;
  int N;
  int S;
  int mode;  // 0: emission, 1: absorption, 2: both spectra of a shared axis
  int n_base, n_base_a;
  float bg;
  const float* __restrict__ vax32;    // [2*ceil(S/2)][2]: (hi, lo) float pair per channel
  const double* __restrict__ basis;   // dataset of the first output (absorption in mode 1)
  const double* __restrict__ offset;
  const double* __restrict__ coeff;
  const double* __restrict__ basis_a; // absorption dataset of mode 2
  const double* __restrict__ offset_a;
  const double* __restrict__ coeff_a;
  const double* __restrict__ cc;
  float* __restrict__ mu;             // [B][S]
  float* __restrict__ mu_a;           // [B][S], mode 2
};
int launch_spectra32(int N, bool pv, const Spec32Args& a, cudaStream_t st);
int launch_spectra(int N, bool pv, const SpecArgs& a, cudaStream_t st);
int launch_spectra_ea(int N, bool pv, const SpecArgs& a, cudaStream_t st);
struct FusedArgs;
int launch_moments_ea(int N, bool pv, const FusedArgs& a, cudaStream_t st);
int launch_moments_em(int N, bool pv, const ChanArgs& a, cudaStream_t st);
int launch_moments_abs(int N, bool pv, const ChanArgs& a, cudaStream_t st);

// cloud count of the instantiation that serves a model with N clouds
inline int padded_clouds(int N) { return N <= 8 ? N : (N <= 12 ? 12 : 16); }

inline unsigned warp_grid(const ChanArgs& a) {
  int64_t items = a.B * a.n_chunks;
  return (unsigned)((items + CHI_WARPS_PER_BLOCK - 1) / CHI_WARPS_PER_BLOCK);
}
}  // namespace chi
