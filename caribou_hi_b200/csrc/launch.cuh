// Launchers of the hot kernels, compiled in their own translation units (launch_*.cu) so that
// the 8 x 2 x 2 template instantiations of each build in parallel.  Return 0, or -1 for an unsupported cloud count.
#pragma once
#include <cuda_runtime.h>

#include "moments.cuh"

namespace chi {
// layout of a caller's theta / grad block (see k_cloud_map in kernels.cuh)
struct RowMap {
  int n_rows;
  signed char row[CHI_NSLOT];
};
// doubles per record of the Jacobian of the parameter maps (k_jacobian, small.cuh)
#define CHI_JAC_W 14

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
// two launches for large Gaussian batches (moments_ea.cuh, MODE): the draws without a negative optical depth
// on `st`, the marked ones on a small persistent grid (launch_ea_tame.cu)
bool moments_ea_two_pass(int N, bool pv, const FusedArgs& a);
int launch_moments_ea_tame(int N, const FusedArgs& a, cudaStream_t st);
int launch_moments_ea_fixup(int N, const FusedArgs& a, int n_sm, cudaStream_t st);
int launch_moments_em(int N, bool pv, const ChanArgs& a, cudaStream_t st);
// two launches for large batches (moments.cuh, MODE; launch_em_tame.cu)
bool moments_em_two_pass(int N, const ChanArgs& a);
int launch_moments_em_tame(int N, bool pv, const ChanArgs& a, cudaStream_t st);
int launch_moments_em_fixup(int N, bool pv, const ChanArgs& a, int n_sm, cudaStream_t st);
int launch_moments_abs(int N, bool pv, const ChanArgs& a, cudaStream_t st);
// ---- argument block of k_eval_small (small.cuh) ----
#define CHI_SM_THREADS 256
#define CHI_SM_WARPS (CHI_SM_THREADS / 32)
#define CHI_SM_CCS 17  // doubles per cloud in shared memory: odd stride, lanes of different clouds hit different banks
#define CHI_SM_MAXREC (1 + 2 * CHI_MAX_BASELINE + CHI_MAX_CLOUDS * 14)

struct SmallArgs {
  int64_t B;
  int N;         // clouds of the model
  int lgG;       // log2 of the lane groups a channel's clouds are spread over (NL * 2^lgG >= N)
  int K;         // tangent directions per cloud (free RVs; 7 for CHI_MODEL_PHYSICS)
  RowMap rm;     // layout of theta and grad
  int prior;     // 1: theta holds unconstrained values, priors and Jacobians are added
  int want_grad;
  int fused, has_e, has_a;
  int S_e, S_a, nb_e, nb_a;
  int n_chunks;          // blocks per draw
  int chunk_e, chunk_a;  // channels per chunk
  int stride;            // doubles per (draw, chunk) record
  int jac_warps;         // last warps of a block that evaluate Jacobian records instead of channels
  double bg;
  const double* __restrict__ chan;    // shared axis: [n_sl][ceil(S/32)][8][32] (moments_ea.cuh)
  const double* __restrict__ chan_e;  // else per dataset [n_sl][ceil(S/32)][4][32] (moments.cuh)
  const double* __restrict__ chan_a;
  const double* __restrict__ basis_e;
  const double* __restrict__ basis_a;
  const double* __restrict__ theta;   // [B][n_rows][N]
  const double* __restrict__ coeff_e;
  const double* __restrict__ coeff_a;
  const int32_t* __restrict__ sl;
  const double* __restrict__ lconst;
  double* __restrict__ mom;           // [B][n_chunks][stride]
  double* __restrict__ jac;           // [B][N][K][CHI_JAC_W]
  double* __restrict__ prior_buf;     // [B][N]
  unsigned* __restrict__ counter;     // [B], zero between launches
  double* __restrict__ logp;
  double* __restrict__ grad;
  double* __restrict__ gce;
  double* __restrict__ gca;
  double bsig_e[CHI_MAX_BASELINE], bsig_a[CHI_MAX_BASELINE];
  double bprior_const;  // sum over the baseline coefficients in use of -log sigma_i - .5 log 2 pi
  long long* stamps;  // development aid (CHI_OPT_SMALL_STAMPS): SM-clock offsets of block 0's phases, else null
};

// one-launch evaluation of a small batch (small.cuh); NL = clouds per thread (1 or 2)
int launch_eval_small(int NL, bool pv, const MapCfg& cfg, const SmallArgs& a, size_t dyn_smem, cudaStream_t st);
// the same evaluation with the blocks of a draw as one thread-block cluster (2 <= n_chunks <= 8)
int launch_eval_small_cl(int NL, bool pv, const MapCfg& cfg, const SmallArgs& a, size_t gb_bytes, cudaStream_t st);
int eval_small_blocks_per_sm(int NL);

// cloud count of the instantiation that serves a model with N clouds
inline int padded_clouds(int N) { return N <= 8 ? N : (N <= 12 ? 12 : 16); }

inline unsigned warp_grid(const ChanArgs& a) {
  int64_t items = a.B * a.n_chunks;
  return (unsigned)((items + CHI_WARPS_PER_BLOCK - 1) / CHI_WARPS_PER_BLOCK);
}
}  // namespace chi
