// CUDA kernels of the caribou_hi likelihood hot path for sm_100a.
//
// Pipeline for one batch of B evaluations (draws x chains x sightlines):
//   k_cloud_map      thread per (draw, cloud): free RVs -> spin temperature ->
//                    line-profile constants cc[B][N][12]             (O(B N))
//   k_moments_em     warp per (draw, channel chunk) of the emission spectrum:
//                    tau -> ordered-cloud radiative transfer -> residual -> reverse
//                    pass through the transfer, all in registers; per-cloud channel
//                    sums ("moments") reduced by warp shuffles       (O(B N S_e))
//   k_moments_abs    same for the absorption spectrum                (O(B N S_a))
//   k_epilogue       thread per (draw, cloud): sums the chunks, contracts the moments
//                    with the Jacobian of the per-cloud map (Dual numbers) (O(B N))
// plus k_spectra (predicted spectra only) and small physics-level kernels.
//
// No tensor cores: the work is exp/FMA on the FP64 pipe.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "cloud_map.cuh"
#include "launch.cuh"

namespace chi {
#define CHI_LOG2E_MAP 1.4426950408889634074

// ---------------------------------------------------------------------------
// k_cloud_map
// ---------------------------------------------------------------------------
// `constrained` != null: theta holds UNCONSTRAINED values (PyMC's value variables); they are
// mapped back through the default transforms first and the constrained values are kept for
// the epilogue.
// Layout of the caller's theta / grad blocks: [B][n_rows][N]; row[s] is the row that holds slot
// s, -1 = the caller does not carry that slot (reads as 0, its gradient is dropped).  The slot
// layout [B][CHI_NSLOT][N] is the identity map; the packed form carries only the free RVs a
// model kind has (chi_logp_grad_rows: fewer bytes over PCIe).
// (struct RowMap: launch.cuh)

// the map of cloud n of one draw: theta_draw / constrained_draw point at the draw's [n_rows][N] /
// [CHI_NSLOT][N] blocks (global or shared memory), cc_draw at its [N][CHI_CC_STRIDE] constants
template <int MODEL>
__device__ __forceinline__ void cloud_map_draw(const MapCfg& cfg, const RowMap& rm, int N,
                                               const double* __restrict__ theta_draw, double* __restrict__ cc_draw,
                                               double* __restrict__ constrained_draw, int n) {
  double th[CHI_NSLOT];
  const double* tp = theta_draw + n;
#pragma unroll
  for (int s = 0; s < CHI_NSLOT; ++s) th[s] = rm.row[s] >= 0 ? tp[(int64_t)rm.row[s] * N] : 0.0;
  if (constrained_draw) {
    double* xp = constrained_draw + n;
#pragma unroll
    for (int s = 0; s < CHI_NSLOT; ++s) {
      th[s] = slot_backward(slot_dist<MODEL>(cfg, s), th[s]).x;
      xp[(int64_t)s * N] = th[s];
    }
  }
  CloudPhys<double> ph = cloud_phys<MODEL, double>(cfg, th);
  double* o = cc_draw + (int64_t)n * CHI_CC_STRIDE;
  o[CC_C] = ph.velocity;
  o[CC_TS] = ph.tspin;
  o[CC_F] = ph.ff;
  ProfConst<double> pe = profile_consts<double>(ph.fwhm2, ph.fwhm_L, ph.tau_total, cfg.wch2_e,
                                                cfg.lorentzian != 0);
  o[CC_KE] = pe.k; o[CC_AGE] = pe.AG; o[CC_ALE] = pe.AL; o[CC_HE] = pe.h;
  // absorption optical depth carries the absorption weight when the model also has
  // emission (emission_absorption_model.py:84-88); absorption-only models have wt == 1
  ProfConst<double> pa = profile_consts<double>(ph.fwhm2, ph.fwhm_L, ph.wt * ph.tau_total,
                                                cfg.wch2_a, cfg.lorentzian != 0);
  o[CC_KA] = pa.k; o[CC_AGA] = pa.AG; o[CC_ALA] = pa.AL; o[CC_HA] = pa.h;
  // 1: an optical depth of this cloud is negative, so 2^y can overflow on it (moments_ea.cuh, MODE); NaNs
  // need no care, they propagate through the remainder of the exponential whatever the branch
  // ... or its peak depth (Gaussian part + Lorentzian part, AL / h at the line centre) is so large that 2^y of it
  // can underflow (summed over the N clouds for the absorption spectrum): the plain forms of the kernels leave
  // out that select as well
  const double big = (1000.0 / CHI_LOG2E_MAP) / N;
  o[CC_CAREFUL] = (pe.AG < 0.0 || pe.AL < 0.0 || pa.AG < 0.0 || pa.AL < 0.0 || pe.AG + pe.AL / pe.h > big ||
                   pa.AG + pa.AL / pa.h > big) ? 1.0 : 0.0;
}

// the map of one (draw, cloud) of a batch in global memory; also called from the sampler's tick kernel
template <int MODEL>
__device__ __forceinline__ void cloud_map_one(const MapCfg& cfg, const RowMap& rm, int N, const double* __restrict__ theta,
                                              double* __restrict__ cc, double* __restrict__ constrained, int64_t b,
                                              int n) {
  cloud_map_draw<MODEL>(cfg, rm, N, theta + (b * rm.n_rows) * N, cc + (b * N) * CHI_CC_STRIDE,
                        constrained ? constrained + (b * CHI_NSLOT) * N : nullptr, n);
}

template <int MODEL>
__global__ void __launch_bounds__(128) k_cloud_map(MapCfg cfg, RowMap rm, int64_t B, int N,
                                                   const double* __restrict__ theta,
                                                   double* __restrict__ cc,
                                                   double* __restrict__ constrained) {
  int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= B * N) return;
  const int64_t b = t / N;
  cloud_map_one<MODEL>(cfg, rm, N, theta, cc, constrained, b, (int)(t - b * N));
}

// ---------------------------------------------------------------------------
// k_epilogue: chunk sums + chain rule through the per-cloud map (Dual numbers)
// ---------------------------------------------------------------------------
struct EpiArgs {
  int64_t B;
  int N;
  RowMap rm;  // layout of theta and grad (identity in model mode)
  int pv;
  int has_e, has_a;
  int nb_e, nb_a;
  int chunks_e, chunks_a;
  int stride_e, stride_a;
  int fused;  // 1: mom_e holds the fused emission+absorption record of moments_ea.cuh
  const double* __restrict__ theta;
  const double* __restrict__ mom_e;
  const double* __restrict__ mom_a;
  const double* __restrict__ lconst;  // [n_sl] sum(-log sigma - 0.5 log 2pi), or null (vjp)
  const int32_t* __restrict__ sl;
  double* __restrict__ logp;          // [B] or null
  double* __restrict__ grad;          // [B][NSLOT][N] or null
  double* __restrict__ gcoeff_e;      // [B][nb_e] or null
  double* __restrict__ gcoeff_a;
  // unconstrained-space model logp (priors + Jacobians), null/0 otherwise
  const double* __restrict__ u;        // [B][NSLOT][N] unconstrained values (theta = constrained)
  const double* __restrict__ coeff_e;  // [B][nb_e] baseline coefficients (their Normal priors)
  const double* __restrict__ coeff_a;
  double bsig_e[CHI_MAX_BASELINE], bsig_a[CHI_MAX_BASELINE];  // prior widths of the coefficients
};

// SPLIT = false: one thread per (draw, cloud) carries all K tangent directions (least work,
// used for large batches).  SPLIT = true: one thread per (draw, cloud, direction) carries a
// single tangent -- ~2x the arithmetic but a ~5x shorter dependent chain, which is what a
// small batch (a few dozen chains) is bound by.
// PRIOR: unconstrained-space model logp (a.u set): prior densities, Jacobians, transform chain rule.
// the large-batch likelihood form runs 4 blocks per SM (128 registers, a few spills): 0.123 -> 0.113 ms at config 2
template <int MODEL, bool SPLIT, bool PRIOR>
__global__ void __launch_bounds__(128, (!SPLIT && !PRIOR) ? 4 : 1) k_epilogue(MapCfg cfg, EpiArgs a) {
  constexpr int K = (MODEL == CHI_MODEL_PHYSICS) ? 7 : CHI_NSLOT;
  constexpr int KD = SPLIT ? 1 : K;   // tangent directions per thread
  constexpr int NJ = K / KD;          // threads per (draw, cloud)
  int64_t b;
  int n, j0;  // cloud, first tangent direction of this thread
  bool valid = true;
  if (PRIOR) {
    // whole draws per block (blockDim is a multiple of N * NJ): the per-cloud prior terms are
    // summed through shared memory in cloud order, so every thread must reach the barrier
    const int per = a.N * NJ, dl = threadIdx.x / per, r = threadIdx.x - dl * per;
    b = (int64_t)blockIdx.x * (blockDim.x / per) + dl;
    n = r / NJ;
    j0 = r - n * NJ;
    valid = b < a.B;
    if (!valid) { b = 0; n = 0; j0 = 0; }
  } else {
    int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= a.B * a.N * NJ) return;
    j0 = SPLIT ? (int)(t % NJ) : 0;
    const int64_t bn = SPLIT ? t / NJ : t;
    b = bn / a.N;
    n = (int)(bn - b * a.N);
  }
  const int NE = a.pv ? 8 : 5, NA = a.pv ? 6 : 3;
  double lp_draw = 0.0;  // likelihood (+ baseline priors) of the draw, held by its first thread

  // --- sums over channel chunks (fixed order: deterministic) ---
  double me[8] = {0, 0, 0, 0, 0, 0, 0, 0}, ma[6] = {0, 0, 0, 0, 0, 0};
  double mf[10] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
  if (a.fused) {
    const int NF = a.pv ? 10 : 6;
    const double* p = a.mom_e + (b * a.chunks_e) * a.stride_e + 1 + a.nb_e + a.nb_a + n * NF;
    for (int c = 0; c < a.chunks_e; ++c, p += a.stride_e)
      for (int j = 0; j < NF; ++j) mf[j] += p[j];
  } else if (a.has_e) {
    const double* p = a.mom_e + (b * a.chunks_e) * a.stride_e + 1 + a.nb_e + n * NE;
    for (int c = 0; c < a.chunks_e; ++c, p += a.stride_e)
      for (int j = 0; j < NE; ++j) me[j] += p[j];
  }
  if (a.has_a && !a.fused) {
    const double* p = a.mom_a + (b * a.chunks_a) * a.stride_a + 1 + a.nb_a + n * NA;
    for (int c = 0; c < a.chunks_a; ++c, p += a.stride_a)
      for (int j = 0; j < NA; ++j) ma[j] += p[j];
  }
  if (n == 0 && j0 == 0 && valid) {
    // logp and baseline-coefficient gradients: one thread per draw
    double lp = 0.0;
    if (a.fused) {
      const double* p = a.mom_e + (b * a.chunks_e) * a.stride_e;
      for (int c = 0; c < a.chunks_e; ++c) lp += p[c * a.stride_e];
      for (int i = 0; i < a.nb_e + a.nb_a; ++i) {
        double g = 0.0;
        for (int c = 0; c < a.chunks_e; ++c) g += p[c * a.stride_e + 1 + i];
        if (i < a.nb_e) { if (a.gcoeff_e) a.gcoeff_e[b * a.nb_e + i] = g; }
        else if (a.gcoeff_a) a.gcoeff_a[b * a.nb_a + (i - a.nb_e)] = g;
      }
    } else if (a.has_e) {
      const double* p = a.mom_e + (b * a.chunks_e) * a.stride_e;
      for (int c = 0; c < a.chunks_e; ++c) lp += p[c * a.stride_e];
      if (a.gcoeff_e)
        for (int i = 0; i < a.nb_e; ++i) {
          double g = 0.0;
          for (int c = 0; c < a.chunks_e; ++c) g += p[c * a.stride_e + 1 + i];
          a.gcoeff_e[b * a.nb_e + i] = g;
        }
    }
    if (a.has_a && !a.fused) {
      const double* p = a.mom_a + (b * a.chunks_a) * a.stride_a;
      for (int c = 0; c < a.chunks_a; ++c) lp += p[c * a.stride_a];
      if (a.gcoeff_a)
        for (int i = 0; i < a.nb_a; ++i) {
          double g = 0.0;
          for (int c = 0; c < a.chunks_a; ++c) g += p[c * a.stride_a + 1 + i];
          a.gcoeff_a[b * a.nb_a + i] = g;
        }
    }
    if (PRIOR) {
      // Normal(0, sigma_i) priors of the baseline coefficients (bayes_spec add_baseline_priors)
      for (int i = 0; i < a.nb_e; ++i) {
        const double c = a.coeff_e ? a.coeff_e[b * a.nb_e + i] : 0.0, s = a.bsig_e[i];
        lp += -0.5 * (c / s) * (c / s) - log(s) - CHI_HALF_LOG_2PI;
        if (a.gcoeff_e) a.gcoeff_e[b * a.nb_e + i] -= c / (s * s);
      }
      for (int i = 0; i < a.nb_a; ++i) {
        const double c = a.coeff_a ? a.coeff_a[b * a.nb_a + i] : 0.0, s = a.bsig_a[i];
        lp += -0.5 * (c / s) * (c / s) - log(s) - CHI_HALF_LOG_2PI;
        if (a.gcoeff_a) a.gcoeff_a[b * a.nb_a + i] -= c / (s * s);
      }
    }
    lp_draw = lp + (a.lconst ? a.lconst[a.sl ? a.sl[b] : 0] : 0.0);
    if (!PRIOR && a.logp) a.logp[b] = lp_draw;
  }
  if (!a.grad && !PRIOR) return;

  // --- Jacobian of the per-cloud map by forward-mode duals ---
  typedef Dual<KD> D;
  D th[CHI_NSLOT];
  const double* tp = a.theta + (b * a.rm.n_rows) * a.N + n;
#pragma unroll
  for (int s = 0; s < CHI_NSLOT; ++s) {
    th[s] = D(a.rm.row[s] >= 0 ? tp[(int64_t)a.rm.row[s] * a.N] : 0.0);
#pragma unroll
    for (int i = 0; i < KD; ++i) th[s].d[i] = (s == j0 + i) ? 1.0 : 0.0;
  }
  CloudPhys<D> ph = cloud_phys<MODEL, D>(cfg, th);

  double g[KD];
#pragma unroll
  for (int i = 0; i < KD; ++i) g[i] = 0.0;
  auto addc = [&](double bar, const D& x) {
#pragma unroll
    for (int i = 0; i < KD; ++i) g[i] = fma(bar, x.d[i], g[i]);
  };

  double cbar = 0.0;
  if (a.fused) {
    // shared axis: k and h are common to the two datasets, their adjoints arrive combined
    const double f = ph.ff.v;
    {
      ProfConst<D> pe = profile_consts<D>(ph.fwhm2, ph.fwhm_L, ph.tau_total, cfg.wch2_e, cfg.lorentzian != 0);
      cbar = 2.0 * pe.k.v * mf[2] + 2.0 * mf[9];
      addc(-mf[3], pe.k);
      addc(f * mf[0], pe.AG);
      if (a.pv) {
        addc(f * mf[6], pe.AL);
        addc(-mf[8], pe.h);
      }
    }
    {
      ProfConst<D> pa = profile_consts<D>(ph.fwhm2, ph.fwhm_L, D(ph.wt * ph.tau_total), cfg.wch2_a,
                                          cfg.lorentzian != 0);
      addc(mf[1], pa.AG);
      if (a.pv) addc(mf[7], pa.AL);
    }
    addc(f * mf[4], ph.tspin);
    addc(mf[5], ph.ff);
  } else if (a.has_e) {
    ProfConst<D> pe = profile_consts<D>(ph.fwhm2, ph.fwhm_L, ph.tau_total, cfg.wch2_e, cfg.lorentzian != 0);
    const double f = ph.ff.v;
    cbar += f * (2.0 * pe.k.v * pe.AG.v * me[1] + 2.0 * pe.AL.v * me[7]);
    addc(-f * pe.AG.v * me[2], pe.k);
    addc(f * me[0], pe.AG);
    if (a.pv) {
      addc(f * me[5], pe.AL);
      addc(-f * pe.AL.v * me[6], pe.h);
    }
    addc(f * me[3], ph.tspin);
    addc(me[4], ph.ff);
  }
  if (a.has_a && !a.fused) {
    ProfConst<D> pa = profile_consts<D>(ph.fwhm2, ph.fwhm_L, D(ph.wt * ph.tau_total), cfg.wch2_a,
                                        cfg.lorentzian != 0);
    cbar += 2.0 * pa.k.v * pa.AG.v * ma[1] + 2.0 * pa.AL.v * ma[5];
    addc(-pa.AG.v * ma[2], pa.k);
    addc(ma[0], pa.AG);
    if (a.pv) {
      addc(ma[3], pa.AL);
      addc(-pa.AL.v * ma[4], pa.h);
    }
  }
  addc(cbar, ph.velocity);

  if (PRIOR) {
    // priors of this cloud's free RVs at their constrained values (Dual: the coupled Normal's
    // mean depends on tkin / fwhm2_thermal), then the chain rule through the transforms
    const double* up = a.u + (b * CHI_NSLOT) * a.N + n;
    const D thermal = (MODEL == CHI_MODEL_EMISSION_ABSORPTION) ? ph.tkin : ph.fwhm2_thermal;
    D lp(0.0);
    double logj = 0.0;
#pragma unroll
    for (int s = 0; s < CHI_NSLOT; ++s) {
      const int dist = slot_dist<MODEL>(cfg, s);
      if (dist == PD_NONE) continue;
      const double w = (s == CHI_FR_FWHM_L_NORM) ? cfg.fwhm_L_weight : 1.0;
      lp = lp + w * slot_density<D>(cfg, dist, th[s], thermal);
      logj += w * slot_backward(dist, up[(int64_t)s * a.N]).logj;
    }
    __shared__ double s_prior[128];
    s_prior[threadIdx.x] = (valid && j0 == 0) ? lp.v + logj : 0.0;
    __syncthreads();
    if (valid && n == 0 && j0 == 0 && a.logp) {
      double tot = lp_draw;
      for (int i = 0; i < a.N; ++i) tot += s_prior[threadIdx.x + i * NJ];
      a.logp[b] = tot;
    }
#pragma unroll
    for (int i = 0; i < KD; ++i) {
      const int s = j0 + i;
      const int dist = slot_dist<MODEL>(cfg, s);
      const double w = (s == CHI_FR_FWHM_L_NORM) ? cfg.fwhm_L_weight : 1.0;
      const Transformed tr = slot_backward(dist, up[(int64_t)s * a.N]);
      g[i] = (g[i] + lp.d[i]) * tr.dxdu + w * tr.dlogj;
    }
  }
  if (!a.grad || !valid) return;

  double* go = a.grad + (b * a.rm.n_rows) * a.N + n;
#pragma unroll
  for (int i = 0; i < KD; ++i)
    if (a.rm.row[j0 + i] >= 0) go[(int64_t)a.rm.row[j0 + i] * a.N] = g[i];
  if (j0 == 0)
    for (int s = K; s < CHI_NSLOT; ++s)  // slots the model kind does not use
      if (a.rm.row[s] >= 0) go[(int64_t)a.rm.row[s] * a.N] = 0.0;
}

// ---------------------------------------------------------------------------
// Small batches (a few dozen chains): the epilogue above is one dependent chain -- chunk sums, then
// the Dual evaluation of the map, then the contraction -- and it sits behind the channel kernels.
// The Jacobian of the map depends on theta only, so it is split off: k_jacobian runs on a side
// stream NEXT TO the channel kernels, and k_contract (a block per draw: coalesced chunk sums into
// shared memory, then one thread per (cloud, direction)) is all that is left on the critical path.
// Same arithmetic as k_epilogue<MODEL, true, PRIOR>; record per (draw, cloud, direction):
//   [dk_e, dAG_e, dAL_e, dh_e, dk_a, dAG_a, dAL_a, dh_a, dTs, df, dc, d(prior density), dx/du, w dlogJ/du]
// ---------------------------------------------------------------------------
struct JacArgs {
  int64_t B;
  int N;
  RowMap rm;
  int has_e, has_a;
  const double* __restrict__ theta;  // constrained values
  const double* __restrict__ u;      // unconstrained values (PRIOR)
  double* __restrict__ jac;          // [B][N][K][CHI_JAC_W]
  double* __restrict__ prior;        // [B][N] prior density + log-Jacobian of the cloud (PRIOR)
};

template <int MODEL, bool PRIOR>
__global__ void __launch_bounds__(128) k_jacobian(MapCfg cfg, JacArgs a) {
  constexpr int K = (MODEL == CHI_MODEL_PHYSICS) ? 7 : CHI_NSLOT;
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= a.B * a.N * K) return;
  const int j0 = (int)(t % K);
  const int64_t bn = t / K;
  const int64_t b = bn / a.N;
  const int n = (int)(bn - b * a.N);
  typedef Dual<1> D;
  D th[CHI_NSLOT];
  const double* tp = a.theta + (b * a.rm.n_rows) * a.N + n;
#pragma unroll
  for (int s = 0; s < CHI_NSLOT; ++s) {
    th[s] = D(a.rm.row[s] >= 0 ? tp[(int64_t)a.rm.row[s] * a.N] : 0.0);
    th[s].d[0] = (s == j0) ? 1.0 : 0.0;
  }
  CloudPhys<D> ph = cloud_phys<MODEL, D>(cfg, th);
  double* o = a.jac + t * CHI_JAC_W;
#pragma unroll
  for (int i = 0; i < CHI_JAC_W; ++i) o[i] = 0.0;
  if (a.has_e) {
    ProfConst<D> pe = profile_consts<D>(ph.fwhm2, ph.fwhm_L, ph.tau_total, cfg.wch2_e, cfg.lorentzian != 0);
    o[0] = pe.k.d[0]; o[1] = pe.AG.d[0]; o[2] = pe.AL.d[0]; o[3] = pe.h.d[0];
  }
  if (a.has_a) {
    ProfConst<D> pa = profile_consts<D>(ph.fwhm2, ph.fwhm_L, D(ph.wt * ph.tau_total), cfg.wch2_a,
                                        cfg.lorentzian != 0);
    o[4] = pa.k.d[0]; o[5] = pa.AG.d[0]; o[6] = pa.AL.d[0]; o[7] = pa.h.d[0];
  }
  o[8] = ph.tspin.d[0]; o[9] = ph.ff.d[0]; o[10] = ph.velocity.d[0];
  if (PRIOR) {
    const double* up = a.u + (b * CHI_NSLOT) * a.N + n;
    const D thermal = (MODEL == CHI_MODEL_EMISSION_ABSORPTION) ? ph.tkin : ph.fwhm2_thermal;
    D lp(0.0);
    double logj = 0.0;
#pragma unroll
    for (int s = 0; s < CHI_NSLOT; ++s) {
      const int dist = slot_dist<MODEL>(cfg, s);
      if (dist == PD_NONE) continue;
      const double w = (s == CHI_FR_FWHM_L_NORM) ? cfg.fwhm_L_weight : 1.0;
      lp = lp + w * slot_density<D>(cfg, dist, th[s], thermal);
      logj += w * slot_backward(dist, up[(int64_t)s * a.N]).logj;
    }
    const int dist = slot_dist<MODEL>(cfg, j0);
    const double w = (j0 == CHI_FR_FWHM_L_NORM) ? cfg.fwhm_L_weight : 1.0;
    const Transformed tr = slot_backward(dist, up[(int64_t)j0 * a.N]);
    o[11] = lp.d[0]; o[12] = tr.dxdu; o[13] = w * tr.dlogj;
    if (j0 == 0) a.prior[bn] = lp.v + logj;
  }
}

struct ConArgs {
  EpiArgs e;
  int K;                            // tangent directions per cloud
  const double* __restrict__ cc;    // [B][N][CHI_CC_STRIDE]
  const double* __restrict__ jac;   // k_jacobian's records (null: no gradient wanted)
  const double* __restrict__ prior; // [B][N] (PRIOR)
};

#define CHI_CON_MAXREC 192
template <bool PRIOR>
__global__ void __launch_bounds__(128) k_contract(ConArgs c) {
  const EpiArgs& a = c.e;
  __shared__ double s_e[CHI_CON_MAXREC], s_a[CHI_CON_MAXREC];
  const int64_t b = blockIdx.x;
  const bool use_e = a.fused || a.has_e, use_a = a.has_a && !a.fused;
  // --- sums over channel chunks (fixed order: deterministic), one record element per thread ---
  if (use_e) {
    const double* p = a.mom_e + (b * a.chunks_e) * a.stride_e;
    for (int t = threadIdx.x; t < a.stride_e; t += blockDim.x) {
      double v = 0.0;
#pragma unroll 4
      for (int ch = 0; ch < a.chunks_e; ++ch) v += p[(int64_t)ch * a.stride_e + t];
      s_e[t] = v;
    }
  }
  if (use_a) {
    const double* p = a.mom_a + (b * a.chunks_a) * a.stride_a;
    for (int t = threadIdx.x; t < a.stride_a; t += blockDim.x) {
      double v = 0.0;
#pragma unroll 4
      for (int ch = 0; ch < a.chunks_a; ++ch) v += p[(int64_t)ch * a.stride_a + t];
      s_a[t] = v;
    }
  }
  __syncthreads();
  // --- baseline-coefficient gradients: thread i handles coefficient i ---
  const int nbe = use_e ? a.nb_e : 0, nba = (a.fused || use_a) ? a.nb_a : 0;
  if ((int)threadIdx.x < nbe + nba) {
    const int i = threadIdx.x;
    if (i < nbe) {
      double g = s_e[1 + i];
      if (PRIOR) g -= (a.coeff_e ? a.coeff_e[b * a.nb_e + i] : 0.0) / (a.bsig_e[i] * a.bsig_e[i]);
      if (a.gcoeff_e) a.gcoeff_e[b * a.nb_e + i] = g;
    } else {
      const int k = i - nbe;
      double g = a.fused ? s_e[1 + a.nb_e + k] : s_a[1 + k];
      if (PRIOR) g -= (a.coeff_a ? a.coeff_a[b * a.nb_a + k] : 0.0) / (a.bsig_a[k] * a.bsig_a[k]);
      if (a.gcoeff_a) a.gcoeff_a[b * a.nb_a + k] = g;
    }
  }
  // --- logp: last warp's first thread (the first threads carry the gradient) ---
  if (threadIdx.x == blockDim.x - 32 && a.logp) {
    double lp = (use_e ? s_e[0] : 0.0) + (use_a ? s_a[0] : 0.0);
    if (PRIOR) {
      for (int i = 0; i < nbe; ++i) {
        const double x = a.coeff_e ? a.coeff_e[b * a.nb_e + i] : 0.0, s = a.bsig_e[i];
        lp += -0.5 * (x / s) * (x / s) - log(s) - CHI_HALF_LOG_2PI;
      }
      for (int i = 0; i < nba; ++i) {
        const double x = a.coeff_a ? a.coeff_a[b * a.nb_a + i] : 0.0, s = a.bsig_a[i];
        lp += -0.5 * (x / s) * (x / s) - log(s) - CHI_HALF_LOG_2PI;
      }
    }
    lp += a.lconst ? a.lconst[a.sl ? a.sl[b] : 0] : 0.0;
    if (PRIOR)
      for (int n = 0; n < a.N; ++n) lp += c.prior[b * a.N + n];
    a.logp[b] = lp;
  }
  if (!a.grad || (int)threadIdx.x >= a.N * c.K) return;

  // --- one thread per (cloud, direction): adjoints of the cloud constants times their tangents ---
  const int n = threadIdx.x / c.K, j0 = threadIdx.x - n * c.K;
  const double* cc = c.cc + (b * a.N + n) * CHI_CC_STRIDE;
  const double* J = c.jac + ((b * a.N + n) * c.K + j0) * CHI_JAC_W;
  const double f = cc[CC_F];
  double g = 0.0, cbar = 0.0;
  if (a.fused) {
    const int NF = a.pv ? 10 : 6;
    const double* mf = s_e + 1 + a.nb_e + a.nb_a + n * NF;
    cbar = 2.0 * cc[CC_KE] * mf[2] + (a.pv ? 2.0 * mf[9] : 0.0);
    g = fma(-mf[3], J[0], g);
    g = fma(f * mf[0], J[1], g);
    if (a.pv) {
      g = fma(f * mf[6], J[2], g);
      g = fma(-mf[8], J[3], g);
    }
    g = fma(mf[1], J[5], g);
    if (a.pv) g = fma(mf[7], J[6], g);
    g = fma(f * mf[4], J[8], g);
    g = fma(mf[5], J[9], g);
  } else if (a.has_e) {
    const int NE = a.pv ? 8 : 5;
    const double* me = s_e + 1 + a.nb_e + n * NE;
    const double AG = cc[CC_AGE], AL = cc[CC_ALE];
    cbar += f * (2.0 * cc[CC_KE] * AG * me[1] + (a.pv ? 2.0 * AL * me[7] : 0.0));
    g = fma(-f * AG * me[2], J[0], g);
    g = fma(f * me[0], J[1], g);
    if (a.pv) {
      g = fma(f * me[5], J[2], g);
      g = fma(-f * AL * me[6], J[3], g);
    }
    g = fma(f * me[3], J[8], g);
    g = fma(me[4], J[9], g);
  }
  if (use_a) {
    const int NA = a.pv ? 6 : 3;
    const double* ma = s_a + 1 + a.nb_a + n * NA;
    const double AG = cc[CC_AGA], AL = cc[CC_ALA];
    cbar += 2.0 * cc[CC_KA] * AG * ma[1] + (a.pv ? 2.0 * AL * ma[5] : 0.0);
    g = fma(-AG * ma[2], J[4], g);
    g = fma(ma[0], J[5], g);
    if (a.pv) {
      g = fma(ma[3], J[6], g);
      g = fma(-AL * ma[4], J[7], g);
    }
  }
  g = fma(cbar, J[10], g);
  if (PRIOR) g = (g + J[11]) * J[12] + J[13];
  double* go = a.grad + (b * a.rm.n_rows) * a.N + n;
  if (a.rm.row[j0] >= 0) go[(int64_t)a.rm.row[j0] * a.N] = g;
  if (j0 == 0)
    for (int s = c.K; s < CHI_NSLOT; ++s)  // slots the model kind does not use
      if (a.rm.row[s] >= 0) go[(int64_t)a.rm.row[s] * a.N] = 0.0;
}

// ---------------------------------------------------------------------------
// Predictive draws of the observed nodes (pm.Normal(mu, sigma = data.noise), emission_model.py:164-169
// and the other add_likelihood bodies; what sample_posterior_predictive / sample_prior_predictive
// return): y = mu + sigma z.  z comes from a COUNTER-BASED generator -- splitmix64's finaliser on
// (seed, global draw index, dataset, channel pair) -- so a sharded or sub-batched job reproduces the
// single-call draws bit for bit and the CPU test restates it in a few lines.  Box-Muller on two
// uniforms per channel pair: 53-bit uniforms and fp64 libm for double outputs, two 24-bit uniforms
// from one hash and fp32 libm for float outputs.
// ---------------------------------------------------------------------------
__host__ __device__ __forceinline__ unsigned long long chi_mix64(unsigned long long z) {
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
  return z ^ (z >> 31);
}
#define CHI_GOLDEN64 0x9E3779B97F4A7C15ULL

template <typename T>
__global__ void __launch_bounds__(256) k_add_noise(int64_t B, int S, T* __restrict__ y,
                                                   const double* __restrict__ chan,  // [n_sl][ceil(S/32)][4][32], plane 3 = 1/sigma^2
                                                   const int32_t* __restrict__ sl, unsigned long long seed,
                                                   int64_t draw_offset, int dataset) {
  const int P = (S + 1) >> 1;
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= B * P) return;
  const int64_t b = t / P;
  const int p = (int)(t - b * P), s0 = 2 * p;
  const unsigned long long g = (unsigned long long)(draw_offset + b);
  const unsigned long long base = chi_mix64(seed + CHI_GOLDEN64 * (2ULL * g + (unsigned long long)dataset + 1ULL));
  const int64_t groups = (S + 31) >> 5;
  const double* w = chan + ((int64_t)(sl ? sl[b] : 0) * groups) * 128 + 96;
  const bool two = s0 + 1 < S;
  const int s1 = two ? s0 + 1 : s0;
  const double w0 = w[(int64_t)(s0 >> 5) * 128 + (s0 & 31)], w1 = w[(int64_t)(s1 >> 5) * 128 + (s1 & 31)];
  T* row = y + b * S;
  if (sizeof(T) == 8) {
    const double sig0 = rsqrt(w0), sig1 = rsqrt(w1);
    const unsigned long long h1 = chi_mix64(base + CHI_GOLDEN64 * (2ULL * p + 1ULL));
    const unsigned long long h2 = chi_mix64(base + CHI_GOLDEN64 * (2ULL * p + 2ULL));
    const double u1 = ((double)(h1 >> 11) + 0.5) * (1.0 / 9007199254740992.0);
    const double u2 = ((double)(h2 >> 11) + 0.5) * (1.0 / 9007199254740992.0);
    const double r = sqrt(-2.0 * log(u1));
    double sn, cs;
    sincospi(2.0 * u2, &sn, &cs);
    row[s0] = (T)((double)row[s0] + sig0 * (r * cs));
    if (two) row[s1] = (T)((double)row[s1] + sig1 * (r * sn));
  } else {
    const unsigned long long h1 = chi_mix64(base + CHI_GOLDEN64 * (2ULL * p + 1ULL));
    const float u1 = ((float)(unsigned)(h1 >> 40) + 0.5f) * (1.0f / 16777216.0f);
    const float u2 = ((float)(unsigned)((h1 >> 16) & 0xFFFFFFu) + 0.5f) * (1.0f / 16777216.0f);
    const float r = sqrtf(-2.0f * logf(u1));  // libm logf: the fast intrinsic loses the relative accuracy near u1 = 1
    float sn, cs;
    __sincosf(6.2831853072f * u2, &sn, &cs);  // MUFU.SIN / MUFU.COS on [0, 2 pi): ~5e-7 absolute
    const float sig0 = rsqrtf((float)w0), sig1 = rsqrtf((float)w1);
    row[s0] = (T)((float)row[s0] + sig0 * (r * cs));
    if (two) row[s1] = (T)((float)row[s1] + sig1 * (r * sn));
  }
}

// unconstrained <-> constrained values of the free RVs (direction = +1: u -> x, -1: x -> u)
template <int MODEL>
__global__ void k_transform(MapCfg cfg, int64_t total, int N, int direction, const double* __restrict__ in,
                            double* __restrict__ out) {
  int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= total) return;
  const int s = (int)((t / N) % CHI_NSLOT);
  const int dist = slot_dist<MODEL>(cfg, s);
  const double v = in[t];
  double r = v;
  if (direction > 0) {
    r = slot_backward(dist, v).x;
  } else if (dist == PD_CHI2 || dist == PD_HALFNORMAL || dist == PD_TRUNCNORMAL) {
    r = log(v);
  } else if (dist == PD_BETA || dist == PD_UNIFORM) {
    r = log(v) - log1p(-v);
  }
  out[t] = r;
}

// ---------------------------------------------------------------------------
// physics-level kernels
// ---------------------------------------------------------------------------
// physics.calc_spin_temp, physics.py:58-106
#ifndef CHI_DEVICE_ONLY
__global__ void k_spin_temp(int64_t n, const double* __restrict__ tk, const double* __restrict__ dens,
                            const double* __restrict__ na, double* __restrict__ out) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = spin_temp<double>(tk[i], dens[i], na[i]);
}
#endif

#ifndef CHI_DEVICE_ONLY
__global__ void k_spin_temp_vjp(int64_t n, const double* __restrict__ tk, const double* __restrict__ dens,
                                const double* __restrict__ na, const double* __restrict__ gbar,
                                double* __restrict__ g_tk, double* __restrict__ g_dens,
                                double* __restrict__ g_na) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  typedef Dual<3> D;
  D r = spin_temp<D>(D::variable(tk[i], 0), D::variable(dens[i], 1), D::variable(na[i], 2));
  const double g = gbar[i];
  g_tk[i] = g * r.d[0];
  g_dens[i] = g * r.d[1];
  g_na[i] = g * r.d[2];
}
#endif

// elementwise helpers of physics.py that the reference evaluates through PyTensor
//   CHI_FN_GAUSSIAN                   physics.gaussian   (x, center, fwhm)        physics.py:16-35
//   CHI_FN_LORENTZIAN                 physics.lorentzian (x, center, fwhm)        physics.py:38-55
//   CHI_FN_LOG10_COLUMN_DENSITY       (tau_total, spin_temp, -)                   physics.py:191-207
//   CHI_FN_LOG10_NONTHERMAL_PRESSURE  (log10_density, fwhm2_nonthermal, -)        physics.py:247-262
#ifndef CHI_DEVICE_ONLY
__global__ void k_elementwise(int fn, int64_t n, const double* __restrict__ a, const double* __restrict__ b,
                              const double* __restrict__ c, double* __restrict__ out) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  double r;
  switch (fn) {
    case CHI_FN_GAUSSIAN: {
      const double d = a[i] - b[i], w2 = c[i] * c[i];
      r = exp(-4.0 * CHI_LN2 * (d * d) / w2) * sqrt(4.0 * CHI_LN2 / (CHI_PI * w2));
    } break;
    case CHI_FN_LORENTZIAN: {
      const double d = a[i] - b[i], hw = c[i] / 2.0;
      r = c[i] / (2.0 * CHI_PI) / (d * d + hw * hw);
    } break;
    case CHI_FN_LOG10_COLUMN_DENSITY:
      r = log10(CHI_COLUMN_CONST * b[i] * a[i]);
      break;
    case CHI_FN_LOG10_NONTHERMAL_PRESSURE:
      r = log10(CHI_PNTH_CONST) + a[i] + log10(b[i]);
      break;
    default:
      r = __longlong_as_double(0x7ff8000000000000LL);
  }
  out[i] = r;
}
#endif

// physics.calc_pseudo_voigt, physics.py:265-309 -> profile[S][N]
#ifndef CHI_DEVICE_ONLY
__global__ void k_pseudo_voigt(int S, int N, const double* __restrict__ vax,
                               const double* __restrict__ velocity, const double* __restrict__ fwhm2,
                               const double* __restrict__ fwhm_L, double* __restrict__ out) {
  int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= (int64_t)S * N) return;
  const int s = (int)(t / N), n = (int)(t % N);
  const double dv = fabs(vax[1] - vax[0]);
  const double wch = 4.0 * CHI_LN2 * dv / CHI_PI;
  ProfConst<double> p = profile_consts<double>(fwhm2[n], fwhm_L[n], 1.0, wch * wch, true);
  const double d = vax[s] - velocity[n], d2 = d * d;
  out[t] = fma(p.AL, 1.0 / (d2 + p.h), p.AG * chi_exp(-p.k * d2));
}
#endif

// physics.radiative_transfer, physics.py:312-365: tau[S][N] -> tb[S]
#ifndef CHI_DEVICE_ONLY
__global__ void k_radiative_transfer(int S, int N, const double* __restrict__ tau,
                                     const double* __restrict__ tspin, const double* __restrict__ ff,
                                     double bg, double* __restrict__ out) {
  int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= S) return;
  double P = 1.0, T = 0.0;
  for (int n = 0; n < N; ++n) {
    const double e = chi_exp(-tau[(int64_t)s * N + n]);
    const double f = ff[n];
    T = fma(f * tspin[n] * (1.0 - e), P, T);
    P *= fma(f, e, 1.0 - f);
  }
  out[s] = (bg * P + T) - bg;
}
#endif

// VJP of physics.calc_pseudo_voigt: gbar[S][N] -> d/d{velocity, fwhm2, fwhm_L}[N].
// One block per cloud; the channel sums are the same six moments the hot kernels use.
#ifndef CHI_DEVICE_ONLY
__global__ void __launch_bounds__(256) k_pseudo_voigt_vjp(int S, int N, const double* __restrict__ vax,
                                                          const double* __restrict__ velocity,
                                                          const double* __restrict__ fwhm2,
                                                          const double* __restrict__ fwhm_L,
                                                          const double* __restrict__ gbar,
                                                          double* __restrict__ g_velocity,
                                                          double* __restrict__ g_fwhm2,
                                                          double* __restrict__ g_fwhm_L) {
  __shared__ double red[8][6];
  const int n = blockIdx.x;
  const double dv = fabs(vax[1] - vax[0]);
  const double wch = 4.0 * CHI_LN2 * dv / CHI_PI;
  const ProfConst<double> p = profile_consts<double>(fwhm2[n], fwhm_L[n], 1.0, wch * wch, true);
  const double c0 = velocity[n];
  double m[6] = {0, 0, 0, 0, 0, 0};
  for (int s = threadIdx.x; s < S; s += blockDim.x) {
    const double g = gbar[(int64_t)s * N + n];
    const double d = vax[s] - c0, d2 = d * d;
    const double E = chi_exp(-p.k * d2), q = 1.0 / (d2 + p.h);
    const double t = g * E, u = g * q, u2 = u * q;
    m[0] += t; m[1] = fma(t, d, m[1]); m[2] = fma(t * d, d, m[2]);
    m[3] += u; m[4] += u2; m[5] = fma(u2, d, m[5]);
  }
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int j = 0; j < 6; ++j) {
    const double r = warp_sum(m[j]);
    if (lane == 0) red[warp][j] = r;
  }
  __syncthreads();
  if (threadIdx.x != 0) return;
  for (int j = 0; j < 6; ++j) {
    double r = 0.0;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) r += red[w][j];
    m[j] = r;
  }
  typedef Dual<2> D;  // tangents: fwhm2, fwhm_L
  const ProfConst<D> pd = profile_consts<D>(D::variable(fwhm2[n], 0), D::variable(fwhm_L[n], 1), D(1.0),
                                            wch * wch, true);
  double g[2] = {0.0, 0.0};
  const double bars[4] = {-p.AG * m[2], m[0], m[3], -p.AL * m[4]};  // kbar, AGbar, ALbar, hbar
  const D* outs[4] = {&pd.k, &pd.AG, &pd.AL, &pd.h};
  for (int j = 0; j < 4; ++j)
    for (int i = 0; i < 2; ++i) g[i] = fma(bars[j], outs[j]->d[i], g[i]);
  g_velocity[n] = 2.0 * p.k * p.AG * m[1] + 2.0 * p.AL * m[5];
  g_fwhm2[n] = g[0];
  g_fwhm_L[n] = g[1];
}
#endif

// VJP of physics.radiative_transfer: gbar[S] -> g_tau[S][N] and (atomically summed over
// channels into zeroed outputs) g_tspin[N], g_ff[N].  N <= CHI_RT_MAX_CLOUDS.
#define CHI_RT_MAX_CLOUDS 64
#ifndef CHI_DEVICE_ONLY
__global__ void __launch_bounds__(128) k_radiative_transfer_vjp(int S, int N, const double* __restrict__ tau,
                                                                const double* __restrict__ tspin,
                                                                const double* __restrict__ ff, double bg,
                                                                const double* __restrict__ gbar,
                                                                double* __restrict__ g_tau,
                                                                double* __restrict__ g_tspin,
                                                                double* __restrict__ g_ff) {
  int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= S) return;
  double om[CHI_RT_MAX_CLOUDS], Pn[CHI_RT_MAX_CLOUDS];
  double P = 1.0;
  for (int n = 0; n < N; ++n) {
    om[n] = 1.0 - chi_exp(-tau[(int64_t)s * N + n]);
    Pn[n] = P;
    P *= fma(-ff[n], om[n], 1.0);
  }
  const double mubar = gbar[s];
  double Q = bg;
  for (int n = N - 1; n >= 0; --n) {
    const double f = ff[n], Ts = tspin[n];
    const double Sn = Ts - Q, x = om[n] * Pn[n];
    g_tau[(int64_t)s * N + n] = mubar * f * Sn * (Pn[n] - x);
    atomicAdd(&g_tspin[n], mubar * f * x);
    atomicAdd(&g_ff[n], mubar * x * Sn);
    Q = fma(fma(-f, om[n], 1.0), Q, f * Ts * om[n]);
  }
}
#endif

// chi_deterministics: every pm.Deterministic of the model class
template <int MODEL>
__global__ void __launch_bounds__(128) k_deterministics(MapCfg cfg, int64_t B, int N,
                                                        const double* __restrict__ theta,
                                                        double* __restrict__ det) {
  int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= B * N) return;
  int64_t b = t / N;
  int n = (int)(t - b * N);
  double th[CHI_NSLOT];
  const double* tp = theta + (b * CHI_NSLOT) * N + n;
#pragma unroll
  for (int s = 0; s < CHI_NSLOT; ++s) th[s] = tp[(int64_t)s * N];
  CloudPhys<double> ph = cloud_phys<MODEL, double, true>(cfg, th);
  const double nan = __longlong_as_double(0x7ff8000000000000LL);
  double o[CHI_NDET];
  for (int i = 0; i < CHI_NDET; ++i) o[i] = nan;
  o[CHI_DET_VELOCITY] = ph.velocity;
  o[CHI_DET_FWHM2] = ph.fwhm2;
  o[CHI_DET_FWHM_L] = ph.fwhm_L;
  o[CHI_DET_TAU_TOTAL] = ph.tau_total;
  if (ph.has_tspin) o[CHI_DET_TSPIN] = ph.tspin;
  if (ph.has_ff) o[CHI_DET_FILLING_FACTOR] = ph.ff;
  if (ph.has_wt) o[CHI_DET_ABSORPTION_WEIGHT] = ph.wt;
  if constexpr (MODEL != CHI_MODEL_PHYSICS) {
    o[CHI_DET_TKIN] = ph.tkin;
    o[CHI_DET_LOG10_nHI] = ph.log10_nHI;
    if (ph.has_tspin) {
      o[CHI_DET_LOG10_NHI] = ph.log10_NHI;
      o[CHI_DET_LOG10_DEPTH] = ph.log10_depth;
    }
    o[CHI_DET_LOG10_N_ALPHA] = ph.log10_n_alpha;
    if (ph.has_TB) o[CHI_DET_TB_FWHM] = ph.TB_fwhm;
    if (ph.has_fraction) o[CHI_DET_FWHM2_THERMAL_FRACTION] = ph.thermal_fraction;
    if (ph.has_thermal) o[CHI_DET_DEPTH] = ph.depth;  // hi_physical_model.py:163
    o[CHI_DET_LOG10_PTH] = log10(ph.tkin) + ph.log10_nHI;  // hi_model.py:143, hi_physical_model.py:202
    if (ph.has_thermal) {
      o[CHI_DET_FWHM2_THERMAL] = ph.fwhm2_thermal;
      o[CHI_DET_FWHM2_NONTHERMAL] = ph.fwhm2_nonthermal;
      // physics.py:262, hi_physical_model.py:206-220
      o[CHI_DET_LOG10_PNTH] = log10(CHI_PNTH_CONST) + ph.log10_nHI + log10(ph.fwhm2_nonthermal);
      o[CHI_DET_LOG10_PTOT] = log10(pow(10.0, o[CHI_DET_LOG10_PTH]) + pow(10.0, o[CHI_DET_LOG10_PNTH]));
    }
  }
  double* op = det + (b * CHI_NDET) * N + n;
  for (int i = 0; i < CHI_NDET; ++i) op[(int64_t)i * N] = o[i];
}

// own-microbenchmark: FP64 FMA peak (8 independent chains per thread)
#ifndef CHI_DEVICE_ONLY
__global__ void __launch_bounds__(256) k_dfma_peak(double* out, int iters, double a, double b) {
  double x0 = threadIdx.x, x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3, x4 = x0 + 4, x5 = x0 + 5, x6 = x0 + 6,
         x7 = x0 + 7;
  for (int i = 0; i < iters; ++i) {
    x0 = fma(x0, a, b); x1 = fma(x1, a, b); x2 = fma(x2, a, b); x3 = fma(x3, a, b);
    x4 = fma(x4, a, b); x5 = fma(x5, a, b); x6 = fma(x6, a, b); x7 = fma(x7, a, b);
  }
  double s = x0 + x1 + x2 + x3 + x4 + x5 + x6 + x7;
  if (s == 12345.678) out[0] = s;  // never true; keeps the chains alive
}
#endif

// FP32 FMA peak (16 independent chains per thread) and SFU peak (MUFU.EX2, 8 chains): the denominators of
// the single-precision spectra kernels' roofline
#ifndef CHI_DEVICE_ONLY
__global__ void __launch_bounds__(256) k_ffma_peak(float* out, int iters, float a, float b) {
  float x[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) x[i] = threadIdx.x + i;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 16; ++i) x[i] = fmaf(x[i], a, b);
  }
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < 16; ++i) s += x[i];
  if (s == 12345.678f) out[0] = s;
}
#endif
#ifndef CHI_DEVICE_ONLY
__global__ void __launch_bounds__(256) k_sfu_peak(float* out, int iters) {
  float x[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) x[i] = -1e-3f * (threadIdx.x + i);
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(x[i]));
  }
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) s += x[i];
  if (s == 12345.678f) out[0] = s;
}
#endif

}  // namespace chi
