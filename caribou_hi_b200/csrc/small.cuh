// One launch per batched logp + gradient evaluation of a SMALL batch (a few NUTS chains: BASELINE
// configs[0] and configs[2]), where the evaluation is bound by dependent latency, not by throughput.
//
// The large-batch pipeline (k_cloud_map -> k_moments_* -> k_epilogue) is three dependent launches
// with one warp per (draw, channel chunk) and every cloud of a draw unrolled inside one thread: at 64
// chains x 8 clouds that is ~1000 long-running warps on a machine with room for 9000, 255 registers
// and spills.  k_eval_small instead
//   * spreads the CLOUDS of a channel over the lanes of a warp ("cloud split": G = 2^lgG lane groups
//     of NL clouds each; a warp iteration covers 32 / G channels).  The ordered-cloud transfer
//     (physics.py:345-365) is a composition of affine maps Q -> A Q + T per cloud (A = 1 - f(1-e),
//     T = f Ts (1-e): attenuate what arrives from behind, add the cloud's own emission), so the
//     attenuation in front of a cloud, the brightness arriving behind it and the emergent spectrum
//     come from an ordered butterfly scan over the lane groups (log2 G shuffle steps of two doubles).
//     A thread carries NL x NF moment accumulators instead of N x NF: no spills at any N;
//   * does the whole evaluation in one kernel: every block maps the free RVs of its draw itself
//     (warp 0; the map is O(N) and redundant work is free here, a dependent launch is not), its last
//     warp(s) evaluate the block's share of the Jacobian of the parameter maps (forward-mode duals)
//     NEXT TO the channel loop, and the block that finishes a draw last (a counter per draw) sums the
//     chunk records in fixed order and contracts them with the Jacobian: logp and gradient leave the
//     kernel, no grid barrier, no co-residency requirement.
// Reference arithmetic: physics.py:265-365, the add_likelihood() bodies, pm.Normal logp and its
// reverse-mode gradient, as in moments.cuh / moments_ea.cuh (same moments, same record layout).
#pragma once
#include <cooperative_groups.h>

#include "launch.cuh"
#include "moments_ea.cuh"

namespace chi {

// ---- the O(N) part, compiled once (not per channel-loop instantiation) -------------------------
// PyMC's default transform of free RV s (slot_backward with the distribution the model kind gives the slot).
// One thread per (cloud, slot) evaluates it ONCE per draw into shared memory (sm_eval), where the map and every
// Jacobian record of the cloud read it: ten exponentials / sigmoids less on the dependent chain of the map, twenty
// less on that of a Jacobian record.
static __device__ __noinline__ Transformed sm_transform(const MapCfg* cfg, int s, double u) {
  int dist = PD_NONE;
  switch (cfg->model) {
#define CASE(M) case M: dist = slot_dist<M>(*cfg, s); break;
    CASE(CHI_MODEL_PHYSICS) CASE(CHI_MODEL_ABSORPTION) CASE(CHI_MODEL_EMISSION) CASE(CHI_MODEL_EMISSION_ABSORPTION)
    CASE(CHI_MODEL_ABSORPTION_PHYSICAL) CASE(CHI_MODEL_EMISSION_PHYSICAL) CASE(CHI_MODEL_EMISSION_ABSORPTION_PHYSICAL)
#undef CASE
  }
  return slot_backward(dist, u);
}

// tr: the cloud's CHI_NSLOT transformed free RVs (prior), or null: tp holds constrained values
template <int MODEL>
__device__ __forceinline__ void sm_map_impl(const MapCfg& cfg, const RowMap& rm, int N, const Transformed* tr,
                                            const double* __restrict__ tp, double* __restrict__ o) {
  double th[CHI_NSLOT];
  if (tr) {
#pragma unroll
    for (int s = 0; s < CHI_NSLOT; ++s) th[s] = tr[s].x;
  } else {
#pragma unroll
    for (int s = 0; s < CHI_NSLOT; ++s) th[s] = rm.row[s] >= 0 ? tp[(int64_t)rm.row[s] * N] : 0.0;
  }
  CloudPhys<double> ph = cloud_phys<MODEL, double>(cfg, th);
  o[CC_C] = ph.velocity;
  o[CC_TS] = ph.tspin;
  o[CC_F] = ph.ff;
  ProfConst<double> pe = profile_consts<double>(ph.fwhm2, ph.fwhm_L, ph.tau_total, cfg.wch2_e, cfg.lorentzian != 0);
  o[CC_KE] = pe.k; o[CC_AGE] = pe.AG; o[CC_ALE] = pe.AL; o[CC_HE] = pe.h;
  ProfConst<double> pa = profile_consts<double>(ph.fwhm2, ph.fwhm_L, ph.wt * ph.tau_total, cfg.wch2_a,
                                                cfg.lorentzian != 0);
  o[CC_KA] = pa.k; o[CC_AGA] = pa.AG; o[CC_ALA] = pa.AL; o[CC_HA] = pa.h;
}

// free RVs of cloud n of one draw (tp = theta of the draw + n) -> the 11 cloud constants (CC_*)
static __device__ __noinline__ void sm_map(const MapCfg* cfg, const RowMap* rm, int N, const Transformed* tr, const double* tp,
                                           double* o) {
  switch (cfg->model) {
#define CASE(M) case M: sm_map_impl<M>(*cfg, *rm, N, tr, tp, o); break;
    CASE(CHI_MODEL_PHYSICS) CASE(CHI_MODEL_ABSORPTION) CASE(CHI_MODEL_EMISSION) CASE(CHI_MODEL_EMISSION_ABSORPTION)
    CASE(CHI_MODEL_ABSORPTION_PHYSICAL) CASE(CHI_MODEL_EMISSION_PHYSICAL) CASE(CHI_MODEL_EMISSION_ABSORPTION_PHYSICAL)
#undef CASE
  }
}

// one record of the Jacobian of the map: tangents along free RV j0 of cloud n (k_jacobian's record)
template <int MODEL>
__device__ __forceinline__ void sm_jac_impl(const MapCfg& cfg, const RowMap& rm, int N, const Transformed* tr, bool has_e,
                                            bool has_a, const double* __restrict__ tp, int j0,
                                            double* __restrict__ o, double* __restrict__ prior_out) {
  typedef Dual<1> D;
  const bool prior = tr != nullptr;
  D th[CHI_NSLOT];
#pragma unroll
  for (int s = 0; s < CHI_NSLOT; ++s) {
    th[s] = D(prior ? tr[s].x : (rm.row[s] >= 0 ? tp[(int64_t)rm.row[s] * N] : 0.0));
    th[s].d[0] = (s == j0) ? 1.0 : 0.0;
  }
  CloudPhys<D> ph = cloud_phys<MODEL, D>(cfg, th);
#pragma unroll
  for (int i = 0; i < CHI_JAC_W; ++i) o[i] = 0.0;
  if (has_e) {
    ProfConst<D> pe = profile_consts<D>(ph.fwhm2, ph.fwhm_L, ph.tau_total, cfg.wch2_e, cfg.lorentzian != 0);
    o[0] = pe.k.d[0]; o[1] = pe.AG.d[0]; o[2] = pe.AL.d[0]; o[3] = pe.h.d[0];
  }
  if (has_a) {
    ProfConst<D> pa = profile_consts<D>(ph.fwhm2, ph.fwhm_L, D(ph.wt * ph.tau_total), cfg.wch2_a, cfg.lorentzian != 0);
    o[4] = pa.k.d[0]; o[5] = pa.AG.d[0]; o[6] = pa.AL.d[0]; o[7] = pa.h.d[0];
  }
  o[8] = ph.tspin.d[0]; o[9] = ph.ff.d[0]; o[10] = ph.velocity.d[0];
  if (prior) {
    const D thermal = (MODEL == CHI_MODEL_EMISSION_ABSORPTION) ? ph.tkin : ph.fwhm2_thermal;
    D lp(0.0);
    double logj = 0.0;
#pragma unroll
    for (int s = 0; s < CHI_NSLOT; ++s) {
      const int dist = slot_dist<MODEL>(cfg, s);
      if (dist == PD_NONE) continue;
      const double w = (s == CHI_FR_FWHM_L_NORM) ? cfg.fwhm_L_weight : 1.0;
      lp = lp + w * slot_density<D>(cfg, dist, th[s], thermal);
      logj += w * tr[s].logj;
    }
    const double w = (j0 == CHI_FR_FWHM_L_NORM) ? cfg.fwhm_L_weight : 1.0;
    o[11] = lp.d[0]; o[12] = tr[j0].dxdu; o[13] = w * tr[j0].dlogj;
    if (j0 == 0) *prior_out = lp.v + logj;
  }
}

static __device__ __noinline__ void sm_jac(const MapCfg* cfg, const RowMap* rm, int N, const Transformed* tr, int has_e, int has_a,
                                    const double* tp, int j0, double* o, double* prior_out) {
  switch (cfg->model) {
#define CASE(M) case M: sm_jac_impl<M>(*cfg, *rm, N, tr, has_e != 0, has_a != 0, tp, j0, o, prior_out); break;
    CASE(CHI_MODEL_PHYSICS) CASE(CHI_MODEL_ABSORPTION) CASE(CHI_MODEL_EMISSION) CASE(CHI_MODEL_EMISSION_ABSORPTION)
    CASE(CHI_MODEL_ABSORPTION_PHYSICAL) CASE(CHI_MODEL_EMISSION_PHYSICAL) CASE(CHI_MODEL_EMISSION_ABSORPTION_PHYSICAL)
#undef CASE
  }
}

// ---- ordered scan of the per-group transfer maps over the lane groups of a channel --------------
// Group g (lane bits below lgG; clouds g NL .. g NL + NL - 1, nearest first) maps the brightness Q
// arriving from behind it to A Q + T.  Returns the attenuation in front of the group (product of
// the A of nearer groups), the brightness arriving behind it, and the emergent brightness of the
// whole line of sight (background included).
__device__ __forceinline__ void sm_scan(int lgG, int lane, double A, double T, double bg, double& front,
                                        double& behind, double& emergent) {
  double bA = A, bB = T;    // composite of the aligned block of groups this lane's group is in
  double sA = 1.0, sB = 0.0;  // composite of the groups behind this one inside that block
  front = 1.0;
#pragma unroll 1
  for (int k = 0; k < lgG; ++k) {
    const int m = 1 << k;
    const double oA = __shfl_xor_sync(CHI_FULL, bA, m), oB = __shfl_xor_sync(CHI_FULL, bB, m);
    const bool far_half = (lane & m) != 0;  // the partner block is nearer to the observer
    // nearer o farther: A = A_near A_far, B = A_near B_far + B_near
    const double nB = far_half ? fma(oA, bB, oB) : fma(bA, oB, bB);
    if (far_half) {
      front *= oA;
    } else {
      sB = fma(sA, oB, sB);
      sA *= oA;
    }
    bA *= oA;
    bB = nB;
  }
  behind = fma(sA, bg, sB);
  emergent = fma(bA, bg, bB);
}

__device__ __forceinline__ double sm_group_sum(int lgG, double x) {
#pragma unroll 1
  for (int k = 0; k < lgG; ++k) x += __shfl_xor_sync(CHI_FULL, x, 1 << k);
  return x;
}
// sum over the channel lanes of a warp (lane bits lgG..4); lanes 0 .. 2^lgG - 1 hold the result
__device__ __forceinline__ double sm_chan_sum(int lgG, double x) {
#pragma unroll 1
  for (int m = 16; m >= (1 << lgG); m >>= 1) x += __shfl_xor_sync(CHI_FULL, x, m);
  return x;
}

// per-thread view of the channel loops
struct SmCtx {
  int lane, g, cl, cpw;  // lane, cloud group, channel within the warp tile, channels per warp tile
  int cw, ncw;           // channel-warp index and count
  int nt;                // threads of the block
  int lgG;
  const ExpCtx* ek;
};

// ---- fused emission + absorption loop (shared axis), moments_ea.cuh's arithmetic ------------------
// sc: shared cloud constants [NL * G][CHI_SM_CCS] in the CC_* layout; acc[NL][NF] as FusedLayout;
// v3 = {sum w r^2 (both datasets), d/d first baseline coefficient emission, absorption}
template <int NL, bool PV>
__device__ __forceinline__ void sm_loop_ea(const SmCtx& x, const SmallArgs& a, const double* __restrict__ sc,
                                           const double* __restrict__ ch, int c0, int c1, double be0, double ba0,
                                           const double* s_beta, double* s_gb, double (&acc)[NL][PV ? 10 : 6],
                                           double (&v3)[3]) {
  // constants of this thread's clouds (registers)
  double c0_[NL], nk[NL], nAGe[NL], fTs[NL], f[NL], nAGa[NL], Ts[NL], fAGe[NL], nALe[NL], nALa[NL], h[NL], fALe[NL];
#pragma unroll
  for (int i = 0; i < NL; ++i) {
    const double* p = sc + (x.g * NL + i) * CHI_SM_CCS;
    const double ff = p[CC_F], ts = p[CC_TS];
    c0_[i] = p[CC_C]; nk[i] = -CHI_LOG2E * p[CC_KE]; nAGe[i] = -CHI_LOG2E * p[CC_AGE]; fTs[i] = ff * ts; f[i] = ff;
    nAGa[i] = -CHI_LOG2E * p[CC_AGA]; Ts[i] = ts; fAGe[i] = ff * nAGe[i];
    nALe[i] = -CHI_LOG2E * p[CC_ALE]; nALa[i] = -CHI_LOG2E * p[CC_ALA]; h[i] = p[CC_HE]; fALe[i] = ff * nALe[i];
  }
  const ExpCtx& ek = *x.ek;
  const int S = a.S_e, nbe = a.nb_e, nba = a.nb_a;
  const int step = x.ncw * x.cpw;
  for (int s0 = c0 + x.cw * x.cpw; s0 < c1; s0 += step) {  // warp-uniform
    const int s = s0 + x.cl;
    const bool valid = s < c1;
    const int sx = valid ? s : c1 - 1;
    const double* __restrict__ cs = ch + (int64_t)(sx >> 5) * 256 + (sx & 31);
    const double v = __ldg(cs), pe0 = __ldg(cs + 32), pa0 = __ldg(cs + 64);
    double base_e = be0 * pe0, base_a = ba0 * pa0;
    _Pragma("unroll 1") for (int i = 1; i < nbe; ++i) base_e = fma(s_beta[i], a.basis_e[(int64_t)i * S + sx], base_e);
    _Pragma("unroll 1") for (int i = 1; i < nba; ++i) base_a = fma(s_beta[CHI_MAX_BASELINE + i], a.basis_a[(int64_t)i * S + sx], base_a);

    double om[NL], Pl[NL], E[NL], q[NL], d[NL];
    double P = 1.0, T = 0.0, xs = 0.0;
#pragma unroll
    for (int i = 0; i < NL; ++i) {
      d[i] = v - c0_[i];
      const double d2 = d[i] * d[i];
      E[i] = chi_exp2_prod<false>(nk[i], d2, ek);
      xs = fma(nAGa[i], E[i], xs);
      if (PV) {
        double t = nAGe[i] * E[i];
        q[i] = chi_rcp(d2 + h[i]);
        t = fma(nALe[i], q[i], t);
        xs = fma(nALa[i], q[i], xs);
        om[i] = 1.0 - chi_exp2<true>(t, ek);
      } else {
        om[i] = 1.0 - chi_exp2_prod<true>(nAGe[i], E[i], ek);
      }
      Pl[i] = P;
      const double u = om[i] * P;
      T = fma(fTs[i], u, T);
      P = fma(-f[i], u, P);
    }
    double front, Q, emergent;
    sm_scan(x.lgG, x.lane, P, T, a.bg, front, Q, emergent);
    xs = sm_group_sum(x.lgG, xs);
    const double mu_e = (emergent - a.bg) + base_e;   // physics.py:362-365
    const double ea = chi_exp2<true>(xs, ek);
    const double mu_a = (1.0 - ea) + base_a;          // absorption_model.py:91
    const double re = __ldg(cs + 128) - mu_e, ra = __ldg(cs + 192) - mu_a;
    const double mbe = valid ? __ldg(cs + 160) * re : 0.0, mba = valid ? __ldg(cs + 224) * ra : 0.0;
    v3[0] = fma(mbe, re, v3[0]);
    v3[0] = fma(mba, ra, v3[0]);
    v3[1] = fma(mbe, pe0, v3[1]);
    v3[2] = fma(mba, pa0, v3[2]);
    if (x.g == 0) {
      _Pragma("unroll 1") for (int i = 1; i < nbe; ++i)
        s_gb[(i - 1) * x.nt] = fma(mbe, a.basis_e[(int64_t)i * S + sx], s_gb[(i - 1) * x.nt]);
      _Pragma("unroll 1") for (int i = 1; i < nba; ++i)
        s_gb[((nbe > 1 ? nbe - 1 : 0) + i - 1) * x.nt] =
            fma(mba, a.basis_a[(int64_t)i * S + sx], s_gb[((nbe > 1 ? nbe - 1 : 0) + i - 1) * x.nt]);
    }
    const double ga = mba * ea;
#pragma unroll
    for (int i = NL - 1; i >= 0; --i) {
      const double Pn = front * Pl[i];
      const double Sn = Ts[i] - Q;
      const double xx = om[i] * Pn;
      const double mx = mbe * xx;
      acc[i][4] += mx;
      acc[i][5] = fma(mx, Sn, acc[i][5]);
      const double ge = (mbe * Sn) * (Pn - xx);
      acc[i][0] = fma(ge, E[i], acc[i][0]);
      acc[i][1] = fma(ga, E[i], acc[i][1]);
      const double gc = -fma(fAGe[i], ge, nAGa[i] * ga);
      const double t = gc * E[i];
      const double td = t * d[i];
      acc[i][2] += td;
      acc[i][3] = fma(td, d[i], acc[i][3]);
      if (PV) {
        acc[i][6] = fma(ge, q[i], acc[i][6]);
        acc[i][7] = fma(ga, q[i], acc[i][7]);
        const double gl = -fma(fALe[i], ge, nALa[i] * ga);
        const double u2 = gl * q[i] * q[i];
        acc[i][8] += u2;
        acc[i][9] = fma(u2, d[i], acc[i][9]);
      }
      Q = fma(om[i], f[i] * Sn, Q);
    }
  }
  // combined moments carry log2(e)
#pragma unroll
  for (int i = 0; i < NL; ++i) {
    acc[i][2] *= CHI_LN2; acc[i][3] *= CHI_LN2;
    if (PV) { acc[i][8] *= CHI_LN2; acc[i][9] *= CHI_LN2; }
  }
}

// ---- emission only (k_moments_em's arithmetic; EmLayout record) -----------------------------------
template <int NL, bool PV>
__device__ __forceinline__ void sm_loop_em(const SmCtx& x, const SmallArgs& a, const double* __restrict__ sc,
                                           const double* __restrict__ ch, int c0, int c1, double be0,
                                           const double* s_beta, double* s_gb, double (&acc)[NL][PV ? 8 : 5],
                                           double (&v3)[3]) {
  double c0_[NL], nk[NL], nAG[NL], fTs[NL], f[NL], Ts[NL], nAL[NL], h[NL];
#pragma unroll
  for (int i = 0; i < NL; ++i) {
    const double* p = sc + (x.g * NL + i) * CHI_SM_CCS;
    const double ff = p[CC_F], ts = p[CC_TS];
    c0_[i] = p[CC_C]; nk[i] = -CHI_LOG2E * p[CC_KE]; nAG[i] = -CHI_LOG2E * p[CC_AGE]; fTs[i] = ff * ts; f[i] = ff;
    Ts[i] = ts; nAL[i] = -CHI_LOG2E * p[CC_ALE]; h[i] = p[CC_HE];
  }
  const ExpCtx& ek = *x.ek;
  const int S = a.S_e, nb = a.nb_e;
  const int step = x.ncw * x.cpw;
  for (int s0 = c0 + x.cw * x.cpw; s0 < c1; s0 += step) {
    const int s = s0 + x.cl;
    const bool valid = s < c1;
    const int sx = valid ? s : c1 - 1;
    const double* __restrict__ cs = ch + (int64_t)(sx >> 5) * 128 + (sx & 31);
    const double v = __ldg(cs), p0 = __ldg(cs + 32);
    double base = be0 * p0;
    _Pragma("unroll 1") for (int i = 1; i < nb; ++i) base = fma(s_beta[i], a.basis_e[(int64_t)i * S + sx], base);
    double om[NL], Pl[NL], E[NL], q[NL], d[NL];
    double P = 1.0, T = 0.0;
#pragma unroll
    for (int i = 0; i < NL; ++i) {
      d[i] = v - c0_[i];
      const double d2 = d[i] * d[i];
      E[i] = chi_exp2_prod<false>(nk[i], d2, ek);
      if (PV) {
        double t = nAG[i] * E[i];
        q[i] = chi_rcp(d2 + h[i]);
        t = fma(nAL[i], q[i], t);
        om[i] = 1.0 - chi_exp2<true>(t, ek);
      } else {
        om[i] = 1.0 - chi_exp2_prod<true>(nAG[i], E[i], ek);
      }
      Pl[i] = P;
      const double u = om[i] * P;
      T = fma(fTs[i], u, T);
      P = fma(-f[i], u, P);
    }
    double front, Q, emergent;
    sm_scan(x.lgG, x.lane, P, T, a.bg, front, Q, emergent);
    const double mu = (emergent - a.bg) + base;
    const double r = __ldg(cs + 64) - mu;
    const double mubar = valid ? __ldg(cs + 96) * r : 0.0;
    v3[0] = fma(mubar, r, v3[0]);
    v3[1] = fma(mubar, p0, v3[1]);
    if (x.g == 0)
      _Pragma("unroll 1") for (int i = 1; i < nb; ++i)
        s_gb[(i - 1) * x.nt] = fma(mubar, a.basis_e[(int64_t)i * S + sx], s_gb[(i - 1) * x.nt]);
#pragma unroll
    for (int i = NL - 1; i >= 0; --i) {
      const double Pn = front * Pl[i];
      const double Sn = Ts[i] - Q;
      const double xx = om[i] * Pn;
      const double mx = mubar * xx;
      acc[i][3] += mx;
      acc[i][4] = fma(mx, Sn, acc[i][4]);
      const double g = (mubar * Sn) * (Pn - xx);
      const double t = g * E[i];
      const double td = t * d[i];
      acc[i][0] += t;
      acc[i][1] += td;
      acc[i][2] = fma(td, d[i], acc[i][2]);
      if (PV) {
        const double u = g * q[i];
        const double u2 = u * q[i];
        acc[i][5] += u;
        acc[i][6] += u2;
        acc[i][7] = fma(u2, d[i], acc[i][7]);
      }
      Q = fma(om[i], f[i] * Sn, Q);
    }
  }
}

// ---- absorption only (k_moments_abs's arithmetic; AbLayout record) -------------------------------
template <int NL, bool PV>
__device__ __forceinline__ void sm_loop_abs(const SmCtx& x, const SmallArgs& a, const double* __restrict__ sc,
                                            const double* __restrict__ ch, int c0, int c1, double ba0,
                                            const double* s_beta, double* s_gb, double (&acc)[NL][PV ? 6 : 3],
                                            double (&v3)[3]) {
  double c0_[NL], nk[NL], nAG[NL], nAL[NL], h[NL];
#pragma unroll
  for (int i = 0; i < NL; ++i) {
    const double* p = sc + (x.g * NL + i) * CHI_SM_CCS;
    c0_[i] = p[CC_C]; nk[i] = -CHI_LOG2E * p[CC_KA]; nAG[i] = -CHI_LOG2E * p[CC_AGA];
    nAL[i] = -CHI_LOG2E * p[CC_ALA]; h[i] = p[CC_HA];
  }
  const ExpCtx& ek = *x.ek;
  const int S = a.S_a, nb = a.nb_a;
  const int step = x.ncw * x.cpw;
  for (int s0 = c0 + x.cw * x.cpw; s0 < c1; s0 += step) {
    const int s = s0 + x.cl;
    const bool valid = s < c1;
    const int sx = valid ? s : c1 - 1;
    const double* __restrict__ cs = ch + (int64_t)(sx >> 5) * 128 + (sx & 31);
    const double v = __ldg(cs), p0 = __ldg(cs + 32);
    double base = ba0 * p0;
    _Pragma("unroll 1") for (int i = 1; i < nb; ++i) base = fma(s_beta[CHI_MAX_BASELINE + i], a.basis_a[(int64_t)i * S + sx], base);
    double E[NL], q[NL], d[NL];
    double xs = 0.0;
#pragma unroll
    for (int i = 0; i < NL; ++i) {
      d[i] = v - c0_[i];
      const double d2 = d[i] * d[i];
      E[i] = chi_exp2_prod<false>(nk[i], d2, ek);
      xs = fma(nAG[i], E[i], xs);
      if (PV) {
        q[i] = chi_rcp(d2 + h[i]);
        xs = fma(nAL[i], q[i], xs);
      }
    }
    xs = sm_group_sum(x.lgG, xs);
    const double ea = chi_exp2<true>(xs, ek);
    const double mu = (1.0 - ea) + base;
    const double r = __ldg(cs + 64) - mu;
    const double mubar = valid ? __ldg(cs + 96) * r : 0.0;
    v3[0] = fma(mubar, r, v3[0]);
    v3[2] = fma(mubar, p0, v3[2]);
    if (x.g == 0)
      _Pragma("unroll 1") for (int i = 1; i < nb; ++i)
        s_gb[((a.nb_e > 1 ? a.nb_e - 1 : 0) + i - 1) * x.nt] =
            fma(mubar, a.basis_a[(int64_t)i * S + sx], s_gb[((a.nb_e > 1 ? a.nb_e - 1 : 0) + i - 1) * x.nt]);
    const double g = mubar * ea;
#pragma unroll
    for (int i = 0; i < NL; ++i) {
      const double t = g * E[i];
      const double td = t * d[i];
      acc[i][0] += t;
      acc[i][1] += td;
      acc[i][2] = fma(td, d[i], acc[i][2]);
      if (PV) {
        const double u = g * q[i];
        const double u2 = u * q[i];
        acc[i][3] += u;
        acc[i][4] += u2;
        acc[i][5] = fma(u2, d[i], acc[i][5]);
      }
    }
  }
}

// warp totals of a thread's NL x M accumulators -> shared s_red[warp][(g NL + i) M + j] (lanes < G write)
template <int NL, int M>
__device__ __forceinline__ void sm_store_acc(const SmCtx& x, double (&acc)[NL][M], double* s_red_warp, int off) {
  // every accumulator goes through the same butterfly over the channel lanes (sm_chan_sum); one loop over the
  // steps with the NL x M independent shuffles of a step in flight together, not NL x M loops one after another
#pragma unroll 1
  for (int m = 16; m >= (1 << x.lgG); m >>= 1) {
#pragma unroll
    for (int i = 0; i < NL; ++i)
#pragma unroll
      for (int j = 0; j < M; ++j) acc[i][j] += __shfl_xor_sync(CHI_FULL, acc[i][j], m);
  }
  if (x.cl == 0) {
#pragma unroll
    for (int i = 0; i < NL; ++i)
#pragma unroll
      for (int j = 0; j < M; ++j) s_red_warp[off + (x.g * NL + i) * M + j] = acc[i][j];
  }
}

// ---------------------------------------------------------------------------------------------------
// the kernel: block = (draw, chunk)
// record per (draw, chunk): [0] = -0.5 sum w r^2, [1..nb_e], [..nb_a] baseline-coefficient gradients,
// then NP x NF fused moments, or NP x NE emission moments followed by NP x NA absorption moments
// (NP = NL 2^lgG clouds, the model's N first)
// ---------------------------------------------------------------------------------------------------
// shared memory of one block of NT threads.  NT = 256: any model (<= CHI_MAX_CLOUDS clouds); NT = 512 (the chain
// kernel's wide form): at most 8 clouds (NL 2^lgG = 8), so that twice the warp rows fit the same space.
template <int NT>
struct SmSmemT {
  static constexpr int RW = 1 + 2 * CHI_MAX_BASELINE + (NT == CHI_SM_THREADS ? CHI_MAX_CLOUDS : 8) * 14 + 2;  // doubles per warp row
  double T[CHI_EXP_TABLE];
  double cc[CHI_MAX_CLOUDS * 2 * CHI_SM_CCS];  // NL 2^lgG <= 32 clouds
  double beta[2 * CHI_MAX_BASELINE];
  double red[NT / 32][RW];
  double sum[CHI_SM_MAXREC + 2];
  Transformed tr[CHI_MAX_CLOUDS * CHI_NSLOT];  // [cloud][slot] (prior mode)
  MapCfg cfg;
  RowMap rm;
  int last;
};
typedef SmSmemT<CHI_SM_THREADS> SmSmem;

// once per block: exponential table and the map's configuration
template <int NT>
__device__ __forceinline__ void sm_init(SmSmemT<NT>& S, const MapCfg& cfg, const RowMap& rm) {
  const int tid = threadIdx.x;
  if (tid == 0) { S.cfg = cfg; S.rm = rm; }
  for (int i = tid; i < CHI_EXP_TABLE; i += NT) S.T[i] = g_exp2_table[CHI_EXP_REPL ? (i >> 4) << 3 : i];
}

// One (draw, chunk) of the evaluation by the whole block; the first statement after sm_init (or after the
// previous call) must be reached by every thread.  tb / ce_b / ca_b: the draw's theta block [n_rows][N] and
// baseline coefficients (global or shared memory; null = no coefficients).
// CHAIN (nuts_chain.cuh: one block evaluates the same chain tick after tick, n_chunks == 1): the block's
// record is the draw's record, so it stays in shared memory -- no record in global memory, no counter.
// SM_CLUSTER: the n_chunks blocks of a draw are one thread-block CLUSTER (launched with that cluster dimension,
// n_chunks <= 8).  Every block pushes its record, and its Jacobian warps their records, straight into the shared
// memory of the cluster's first block (distributed shared memory: `xbuf` = this block's exchange area
// [n_chunks][stride] records, [N K][CHI_JAC_W] Jacobian records, [N] prior terms); one cluster barrier later the
// first block has everything on its own SM.  This replaces the trip through global memory of SM_GLOBAL -- write
// the record, fence, count the draw's blocks with an atomic, and whichever block comes last reads everything
// back through L2 -- which is ~5 us of a 27 us evaluation at configs[2]'s shape.  Same sums in the same order.
enum { SM_GLOBAL = 0, SM_CHAIN = 1, SM_CLUSTER = 2 };

// SUB: the NT evaluating threads are the first NT threads of a LARGER block (k_nuts_chain_spec: one more warp runs
// the chain's state machine next to them), so their block barrier is a named one over NT threads.
template <bool SUB, int NT>
__device__ __forceinline__ void sm_sync() {
  if (SUB) asm volatile("bar.sync 2, %0;" ::"n"(NT) : "memory");
  else __syncthreads();
}

template <int NL, bool PV, int MODE, int NT = CHI_SM_THREADS, bool SUB = false>
__device__ __forceinline__ void sm_eval(SmSmemT<NT>& S, double* s_gb_dyn, const SmallArgs& a, const int64_t b, const int chunk,
                                        const double* tb, const double* ce_b, const double* ca_b, double* xbuf = nullptr) {
  constexpr bool CHAIN = MODE == SM_CHAIN, CLUSTER = MODE == SM_CLUSTER;
  static_assert(!SUB || CHAIN, "a sub-block evaluation is the chain kernel's");
  constexpr int NF = PV ? 10 : 6, NE = PV ? 8 : 5, NA = PV ? 6 : 3;
  double* const s_T = S.T;
  double* const s_cc = S.cc;
  double* const s_beta = S.beta;
  double(*const s_red)[SmSmemT<NT>::RW] = S.red;
  double* const s_sum = S.sum;
  MapCfg& s_cfg = S.cfg;
  RowMap& s_rm = S.rm;
  int& s_last = S.last;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int N = a.N, G = 1 << a.lgG, NP = NL * G;
  const int nbe = a.nb_e, nba = a.nb_a;

  // (every block of the cluster must be running before anything is stored into the first block's shared memory:
  // arrive here, wait just before the first such store -- the wait hides behind the set-up below)
  if (CLUSTER) cooperative_groups::this_cluster().barrier_arrive();
  // null clouds pad the lane groups: tau = 0, f = 0
  for (int i = tid; i < NP * CHI_SM_CCS; i += NT) {
    const int k = i % CHI_SM_CCS;
    s_cc[i] = (k == CC_KE || k == CC_KA || k == CC_HE || k == CC_HA) ? 1.0 : 0.0;
  }
  if (tid < 2 * CHI_MAX_BASELINE) {
    const bool e = tid < CHI_MAX_BASELINE;
    const int i = e ? tid : tid - CHI_MAX_BASELINE;
    const double* c = e ? ce_b : ca_b;
    const int nb = e ? nbe : nba;
    s_beta[tid] = (c && i < nb) ? c[i] : 0.0;
  }
  const int n_gb = (nbe > 1 ? nbe - 1 : 0) + (nba > 1 ? nba - 1 : 0);
  for (int i = 0; i < n_gb; ++i) s_gb_dyn[i * NT + tid] = 0.0;
  sm_sync<SUB, NT>();
  if (a.prior) {
    // the transforms of the draw's free RVs, one thread each (the channel warps have nothing else to do yet)
    if (tid < N * CHI_NSLOT) {
      const int n = tid / CHI_NSLOT, sl_ = tid - n * CHI_NSLOT;
      const double u = s_rm.row[sl_] >= 0 ? tb[(int64_t)s_rm.row[sl_] * N + n] : 0.0;
      S.tr[tid] = sm_transform(&s_cfg, sl_, u);
    }
    sm_sync<SUB, NT>();
  }

  // where the draw's records meet: the first block of the cluster, or the workspace in global memory
  double* lead = nullptr;
  if (CLUSTER) {
    cooperative_groups::this_cluster().barrier_wait();
    lead = cooperative_groups::this_cluster().map_shared_rank(xbuf, 0);
  }
  double* const jac_b = CLUSTER ? lead + a.n_chunks * a.stride : a.jac + b * N * a.K * CHI_JAC_W;
  double* const prior_b = CLUSTER ? jac_b + N * a.K * CHI_JAC_W : a.prior_buf + b * N;
  // ---- map (warp 0) next to the Jacobian records of this block (last warps) ----
  const int ncw = (NT / 32) - a.jac_warps;  // channel warps
  // the draw's likelihood constant, fetched early by the thread that writes logp at the very end
  double lconst_b = 0.0;
  if (tid == NT - 32 && a.lconst) lconst_b = __ldg(a.lconst + (a.sl ? a.sl[b] : 0));
  long long t_0 = 0, t_map = 0, t_jac = 0, t_chan = 0;
  if (a.stamps) t_0 = clock64();
  if (warp == 0 && lane < N)
    sm_map(&s_cfg, &s_rm, N, a.prior ? S.tr + lane * CHI_NSLOT : nullptr, tb + lane, s_cc + lane * CHI_SM_CCS);
  if (a.stamps) t_map = clock64();
  if (warp >= ncw) {
    // records r = n K + j of this draw with r % n_chunks == chunk
    const int t = tid - ncw * 32;
    const int r = chunk + t * a.n_chunks;
    if (r < N * a.K) {
      const int n = r / a.K, j0 = r - n * a.K;
      if (a.want_grad || (a.prior && j0 == 0))
        sm_jac(&s_cfg, &s_rm, N, a.prior ? S.tr + n * CHI_NSLOT : nullptr, a.fused || a.has_e, a.fused || a.has_a, tb + n, j0,
               jac_b + (n * a.K + j0) * CHI_JAC_W, prior_b + n);
    }
    if (a.stamps) t_jac = clock64();
  } else {
    // the channel warps wait for the map only (named barrier 1); the Jacobian warps run next to them
    asm volatile("bar.sync 1, %0;" ::"r"(ncw * 32) : "memory");
  }

  // ---- channel loops (warps 0 .. ncw-1) ----
  const int stride = a.stride;
  double* red = s_red[warp];
  if (warp < ncw) {
    const ExpCtx ek = CHI_LOAD_EXPCTX(s_T);
    SmCtx x;
    x.lane = lane; x.lgG = a.lgG; x.g = lane & (G - 1); x.cl = lane >> a.lgG; x.cpw = 32 >> a.lgG;
    x.cw = warp; x.ncw = ncw; x.nt = NT; x.ek = &ek;
    const int sl = a.sl ? a.sl[b] : 0;
    double v3[3] = {0.0, 0.0, 0.0};
    double* gb = s_gb_dyn + tid;
    if (a.fused) {
      double acc[NL][NF];
#pragma unroll
      for (int i = 0; i < NL; ++i)
#pragma unroll
        for (int j = 0; j < NF; ++j) acc[i][j] = 0.0;
      const int c0 = chunk * a.chunk_e, c1 = min(a.S_e, c0 + a.chunk_e);
      const double* ch = a.chan + (int64_t)sl * ((a.S_e + 31) >> 5) * 256;
      if (c0 < c1) sm_loop_ea<NL, PV>(x, a, s_cc, ch, c0, c1, s_beta[0], s_beta[CHI_MAX_BASELINE], s_beta, gb, acc, v3);
      sm_store_acc<NL, NF>(x, acc, red, 1 + nbe + nba);
    } else {
      if (a.has_e) {
        double acc[NL][NE];
#pragma unroll
        for (int i = 0; i < NL; ++i)
#pragma unroll
          for (int j = 0; j < NE; ++j) acc[i][j] = 0.0;
        const int c0 = chunk * a.chunk_e, c1 = min(a.S_e, c0 + a.chunk_e);
        const double* ch = a.chan_e + (int64_t)sl * ((a.S_e + 31) >> 5) * 128;
        if (c0 < c1) sm_loop_em<NL, PV>(x, a, s_cc, ch, c0, c1, s_beta[0], s_beta, gb, acc, v3);
        if (a.stamps && b == 0 && chunk == 0 && tid == 0) a.stamps[7] = clock64() - t_0;
        sm_store_acc<NL, NE>(x, acc, red, 1 + nbe + nba);
        if (a.stamps && b == 0 && chunk == 0 && tid == 0) a.stamps[15] = clock64() - t_0;
      }
      if (a.has_a) {
        double acc[NL][NA];
#pragma unroll
        for (int i = 0; i < NL; ++i)
#pragma unroll
          for (int j = 0; j < NA; ++j) acc[i][j] = 0.0;
        const int c0 = chunk * a.chunk_a, c1 = min(a.S_a, c0 + a.chunk_a);
        const double* ch = a.chan_a + (int64_t)sl * ((a.S_a + 31) >> 5) * 128;
        if (c0 < c1) sm_loop_abs<NL, PV>(x, a, s_cc, ch, c0, c1, s_beta[CHI_MAX_BASELINE], s_beta, gb, acc, v3);
        sm_store_acc<NL, NA>(x, acc, red, 1 + nbe + nba + (a.has_e ? NP * NE : 0));
      }
    }
    // logp and the baseline terms live on the lanes of cloud group 0 (every group computed the same values)
#pragma unroll 1
    for (int m = 16; m >= (1 << a.lgG); m >>= 1) {  // the three sums through one butterfly (sm_chan_sum's steps)
#pragma unroll
      for (int j = 0; j < 3; ++j) v3[j] += __shfl_xor_sync(CHI_FULL, v3[j], m);
    }
    const double lp = v3[0], g0e = v3[1], g0a = v3[2];
    if (lane == 0) {
      red[0] = -0.5 * lp;
      if (nbe > 0) red[1] = g0e;
      if (nba > 0) red[1 + nbe] = g0a;
    }
    _Pragma("unroll 1") for (int i = 1; i < nbe; ++i) {
      const double t = sm_chan_sum(a.lgG, x.g == 0 ? gb[(i - 1) * NT] : 0.0);
      if (lane == 0) red[1 + i] = t;
    }
    _Pragma("unroll 1") for (int i = 1; i < nba; ++i) {
      const double t = sm_chan_sum(a.lgG, x.g == 0 ? gb[((nbe > 1 ? nbe - 1 : 0) + i - 1) * NT] : 0.0);
      if (lane == 0) red[1 + nbe + i] = t;
    }
  }
  if (a.stamps) t_chan = clock64();
  sm_sync<SUB, NT>();
  if (a.stamps && b == 0 && chunk == 0 && lane == 0 && (warp == 0 || warp == (NT / 32) - 1)) {
    long long* o = a.stamps + (warp == 0 ? 0 : 8);
    o[0] = t_map - t_0; o[1] = t_jac - t_0; o[2] = t_chan - t_0; o[3] = clock64() - t_0;
  }
  if (CHAIN) {
    // the block's record is the draw's record (0 + v, as the general form sums a single chunk)
    for (int t = tid; t < stride; t += NT) {
      double v = 0.0;
      for (int w = 0; w < ncw; ++w) v += s_red[w][t];
      s_sum[t] = 0.0 + v;
    }
  } else if (CLUSTER) {
    for (int t = tid; t < stride; t += NT) {
      double v = 0.0;
      for (int w = 0; w < ncw; ++w) v += s_red[w][t];
      lead[chunk * stride + t] = v;
    }
    cooperative_groups::this_cluster().sync();  // every record and Jacobian record of the draw is in the first block
    if (a.stamps && b == 0 && chunk == 0 && tid == 0) a.stamps[4] = clock64() - t_0;
    if (chunk != 0) return;
    for (int t = tid; t < stride; t += NT) {
      double v = 0.0;
      for (int c = 0; c < a.n_chunks; ++c) v += xbuf[c * stride + t];  // chunk order, as SM_GLOBAL
      s_sum[t] = v;
    }
  } else {
    // channel-warp partials in fixed order -> the (draw, chunk) record
    double* rec = a.mom + (b * a.n_chunks + chunk) * stride;
    for (int t = tid; t < stride; t += NT) {
      double v = 0.0;
      for (int w = 0; w < ncw; ++w) v += s_red[w][t];
      rec[t] = v;
    }
    __threadfence();
    sm_sync<SUB, NT>();
    if (tid == 0) {
      const unsigned done = atomicAdd(a.counter + b, 1u);
      s_last = (done == (unsigned)a.n_chunks - 1u);
      if (s_last) a.counter[b] = 0;  // ready for the next launch
    }
    sm_sync<SUB, NT>();
    if (a.stamps && b == 0 && chunk == 0 && tid == 0) a.stamps[4] = clock64() - t_0;
    if (!s_last) return;
    __threadfence();

    // ---- this block completed the draw: chunk sums in fixed order, then the contraction ----
    const double* rec0 = a.mom + (b * a.n_chunks) * stride;
    for (int t = tid; t < stride; t += NT) {
      double v = 0.0;
      int c = 0;
      for (; c + 4 <= a.n_chunks; c += 4) {  // four loads in flight, summed in chunk order
        const double* p = rec0 + (int64_t)c * stride + t;
        const double v0 = __ldcg(p), v1 = __ldcg(p + stride), v2 = __ldcg(p + 2 * (int64_t)stride),
                     v3 = __ldcg(p + 3 * (int64_t)stride);
        v = (((v + v0) + v1) + v2) + v3;
      }
      for (; c < a.n_chunks; ++c) v += __ldcg(rec0 + (int64_t)c * stride + t);
      s_sum[t] = v;
    }
  }
  // the per-cloud prior terms of the draw (written by the Jacobian warps of its blocks)
  if (a.prior && tid < N) s_red[0][tid] = CLUSTER ? prior_b[tid] : __ldcg(prior_b + tid);
  sm_sync<SUB, NT>();
  const bool use_e = a.fused || a.has_e, use_a = a.fused || a.has_a;
  const int nbe_u = use_e ? nbe : 0, nba_u = use_a ? nba : 0;
  if (tid < nbe + nba) {
    // baseline-coefficient gradients (zero for a dataset that did not run)
    if (tid < nbe) {
      double g = use_e ? s_sum[1 + tid] : 0.0;
      if (a.prior) g -= s_beta[tid] / (a.bsig_e[tid] * a.bsig_e[tid]);
      if (a.gce) a.gce[b * nbe + tid] = g;
    } else {
      const int k = tid - nbe;
      double g = use_a ? s_sum[1 + nbe + k] : 0.0;
      if (a.prior) g -= s_beta[CHI_MAX_BASELINE + k] / (a.bsig_a[k] * a.bsig_a[k]);
      if (a.gca) a.gca[b * nba + k] = g;
    }
  }
  if (tid == NT - 32 && a.logp) {
    double lp = s_sum[0];
    if (a.prior) {
      // Normal(0, sigma_i) priors of the baseline coefficients; bprior_const = sum(-log sigma_i - .5 log 2 pi)
      for (int i = 0; i < nbe_u; ++i) {
        const double z = s_beta[i] / a.bsig_e[i];
        lp = fma(-0.5 * z, z, lp);
      }
      for (int i = 0; i < nba_u; ++i) {
        const double z = s_beta[CHI_MAX_BASELINE + i] / a.bsig_a[i];
        lp = fma(-0.5 * z, z, lp);
      }
      lp += a.bprior_const;
    }
    lp += lconst_b;
    if (a.prior)
      for (int n = 0; n < N; ++n) lp += s_red[0][n];
    a.logp[b] = lp;
  }
  if (a.stamps && tid == NT - 32) a.stamps[5] = clock64() - t_0;
  if (!a.want_grad || !a.grad || tid >= N * a.K) return;

  const int n = tid / a.K, j0 = tid - n * a.K;
  const double* cc = s_cc + n * CHI_SM_CCS;
  const double* Jg = jac_b + (n * a.K + j0) * CHI_JAC_W;
  double J[CHI_JAC_W];
#pragma unroll
  for (int i = 0; i < CHI_JAC_W; ++i) J[i] = CLUSTER ? Jg[i] : __ldcg(Jg + i);
  const double f = cc[CC_F];
  double g = 0.0, cbar = 0.0;
  const double* m0 = s_sum + 1 + nbe + nba;
  if (a.fused) {
    const double* mf = m0 + n * NF;
    cbar = 2.0 * cc[CC_KE] * mf[2] + (PV ? 2.0 * mf[9] : 0.0);
    g = fma(-mf[3], J[0], g);
    g = fma(f * mf[0], J[1], g);
    if (PV) {
      g = fma(f * mf[6], J[2], g);
      g = fma(-mf[8], J[3], g);
    }
    g = fma(mf[1], J[5], g);
    if (PV) g = fma(mf[7], J[6], g);
    g = fma(f * mf[4], J[8], g);
    g = fma(mf[5], J[9], g);
  } else {
    if (a.has_e) {
      const double* me = m0 + n * NE;
      const double AG = cc[CC_AGE], AL = cc[CC_ALE];
      cbar += f * (2.0 * cc[CC_KE] * AG * me[1] + (PV ? 2.0 * AL * me[7] : 0.0));
      g = fma(-f * AG * me[2], J[0], g);
      g = fma(f * me[0], J[1], g);
      if (PV) {
        g = fma(f * me[5], J[2], g);
        g = fma(-f * AL * me[6], J[3], g);
      }
      g = fma(f * me[3], J[8], g);
      g = fma(me[4], J[9], g);
    }
    if (a.has_a) {
      const double* ma = m0 + (a.has_e ? NP * NE : 0) + n * NA;
      const double AG = cc[CC_AGA], AL = cc[CC_ALA];
      cbar += 2.0 * cc[CC_KA] * AG * ma[1] + (PV ? 2.0 * AL * ma[5] : 0.0);
      g = fma(-AG * ma[2], J[4], g);
      g = fma(ma[0], J[5], g);
      if (PV) {
        g = fma(ma[3], J[6], g);
        g = fma(-AL * ma[4], J[7], g);
      }
    }
  }
  g = fma(cbar, J[10], g);
  if (a.prior) g = (g + J[11]) * J[12] + J[13];
  double* go = a.grad + (b * a.rm.n_rows) * N + n;
  if (a.rm.row[j0] >= 0) go[(int64_t)a.rm.row[j0] * N] = g;
  if (j0 == 0)
    for (int s = a.K; s < CHI_NSLOT; ++s)
      if (a.rm.row[s] >= 0) go[(int64_t)a.rm.row[s] * N] = 0.0;
  if (a.stamps && tid == 0) a.stamps[6] = clock64() - t_0;
}

// Static + dynamic shared memory of these kernels passes the 48 KB a launch gets by default once the baselines
// have several terms (one column of CHI_SM_THREADS doubles per extra term): the opt-in is raised when needed.
template <class Kernel>
inline int sm_allow_dyn(Kernel kernel, size_t dyn_smem) {
  if (dyn_smem <= 2048) return 0;
  return cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dyn_smem) != cudaSuccess;
}

template <int NL, bool PV>
__global__ void __launch_bounds__(CHI_SM_THREADS, NL <= 2 ? 2 : 1) k_eval_small(MapCfg cfg, SmallArgs a) {
  __shared__ SmSmem S;
  extern __shared__ double s_gb_dyn[];  // (nb_e - 1 + nb_a - 1) x CHI_SM_THREADS when baselines have > 1 term
  const int64_t b = blockIdx.x / a.n_chunks;
  const int chunk = (int)(blockIdx.x - b * a.n_chunks);
  sm_init(S, cfg, a.rm);
  sm_eval<NL, PV, SM_GLOBAL>(S, s_gb_dyn, a, b, chunk, a.theta + (b * a.rm.n_rows) * a.N,
                             a.coeff_e ? a.coeff_e + b * a.nb_e : nullptr, a.coeff_a ? a.coeff_a + b * a.nb_a : nullptr);
}

// the cluster form (SM_CLUSTER): launched with cluster dimension n_chunks; x_off = byte offset of the exchange
// area in the dynamic shared memory
template <int NL, bool PV>
__global__ void __launch_bounds__(CHI_SM_THREADS, NL <= 2 ? 2 : 1) k_eval_small_cl(MapCfg cfg, SmallArgs a, unsigned x_off) {
  __shared__ SmSmem S;
  extern __shared__ double s_gb_dyn[];
  const int64_t b = blockIdx.x / a.n_chunks;
  const int chunk = (int)(blockIdx.x - b * a.n_chunks);  // == the block's rank in its cluster
  sm_init(S, cfg, a.rm);
  sm_eval<NL, PV, SM_CLUSTER>(S, s_gb_dyn, a, b, chunk, a.theta + (b * a.rm.n_rows) * a.N,
                              a.coeff_e ? a.coeff_e + b * a.nb_e : nullptr, a.coeff_a ? a.coeff_a + b * a.nb_a : nullptr,
                              reinterpret_cast<double*>(reinterpret_cast<char*>(s_gb_dyn) + x_off));
}

}  // namespace chi
