// Per-cloud O(N) arithmetic of the hot path, templated on the scalar type so the
// same source gives values (double) and Jacobians (Dual<K>):
//   * spin temperature                      reference physics.py:58-106
//   * the six model classes' add_priors()/add_deterministics() maps
//   * the per-dataset line-profile constants of physics.py:294-309
#pragma once
#include "../../include/chi.h"
#include "dual.cuh"

namespace chi {

// constants exactly as the reference writes them
#define CHI_T_RADIATION 3.77        /* physics.py:75 */
#define CHI_LYA_COEFF 5.9e11        /* physics.py:79 */
#define CHI_T0 0.0681               /* physics.py:82 */
#define CHI_A10 2.8843e-15          /* physics.py:85 */
#define CHI_THERMAL_CONST 0.2139    /* physics.py:123,142 */
#define CHI_COLUMN_CONST 1.82243e18 /* physics.py:206,243; emission_physical_model.py:65 */
#define CHI_PC_CM_LOG10 18.489351   /* physics.py:225 */
#define CHI_PNTH_CONST 671.8        /* physics.py:262 */
#define CHI_LN2 0.693147180559945309417
#define CHI_PI 3.14159265358979323846

// Device-side copy of the configuration the maps need.
struct MapCfg {
  int model, lorentzian, free_weight, has_e, has_a;
  double bg_temp, inv_power;
  double prior_fwhm2, pv0, pv1, pn0, pn1, pa0, pa1, prior_fwhm_L, prior_amp;
  double wch2_e, wch2_a;  // (4 ln2 |v1-v0| / pi)^2 per dataset, physics.py:294-295
  double const_G;         // sqrt(2 pi)/(2 sqrt(2 ln 2)), emission_model.py:127 (host-computed)
  // prior terms of the unconstrained-space model logp (chi_set_priors; host-computed constants)
  double beta_a, beta_b, beta_lnB;      // Beta(a, b) of the thermal fraction / tkin factor
  double nth_mu, nth_sigma, nth_const;  // TruncatedNormal(mu, sigma, lower=0): -ln sigma - .5 ln 2pi - ln(1-Phi)
  double wt_sigma, wt_const, wt_mu0;    // coupled Normal: mu = wt_mu0 - log10(T), const = -ln sigma - .5 ln 2pi
  double fwhm_L_weight;                 // 1/N when fwhm_L_norm is one RV per draw (physical models), else 1
};

// ---- priors and default transforms of PyMC's unconstrained space ------
// Distribution of the free RV in each theta slot; the transform follows from it:
//   CHI2 / HALFNORMAL / TRUNCNORMAL(lower=0): x = exp(u), log|J| = u
//   BETA / UNIFORM(0,1):                      x = sigmoid(u), log|J| = log x + log(1-x)
enum { PD_NONE = 0, PD_NORMAL01, PD_CHI2, PD_HALFNORMAL, PD_BETA, PD_UNIFORM, PD_TRUNCNORMAL, PD_COUPLED };

template <int MODEL>
CHI_HD int slot_dist(const MapCfg& c, int s) {
  if (MODEL == CHI_MODEL_PHYSICS) return PD_NONE;
  constexpr bool physical = (MODEL >= CHI_MODEL_ABSORPTION_PHYSICAL);
  constexpr bool emission = (MODEL == CHI_MODEL_EMISSION || MODEL == CHI_MODEL_EMISSION_ABSORPTION ||
                             MODEL == CHI_MODEL_EMISSION_PHYSICAL || MODEL == CHI_MODEL_EMISSION_ABSORPTION_PHYSICAL);
  switch (s) {
    case CHI_FR_FWHM2_NORM: return PD_CHI2;                                   // hi_model.py:95
    case CHI_FR_VELOCITY_NORM: return PD_NORMAL01;                            // hi_model.py:109
    case CHI_FR_LOG10_NHI_NORM: return physical ? PD_NONE : PD_NORMAL01;      // hi_model.py:99
    case CHI_FR_LOG10_N_ALPHA_NORM: return PD_NORMAL01;                       // hi_model.py:122
    case CHI_FR_NTH_FWHM_1PC: return physical ? PD_TRUNCNORMAL : PD_NONE;     // hi_physical_model.py:138
    case CHI_FR_FWHM_L_NORM: return c.lorentzian ? PD_HALFNORMAL : PD_NONE;   // hi_model.py:133
    case CHI_FR_AMPLITUDE_NORM: return PD_HALFNORMAL;                         // e.g. emission_model.py:62
    case CHI_FR_THERMAL: return PD_BETA;                                      // e.g. emission_model.py:76
    case CHI_FR_FILLING_FACTOR_NORM: return emission ? PD_UNIFORM : PD_NONE;  // emission_model.py:107
    case CHI_FR_LOG10_WT:
      if (MODEL == CHI_MODEL_EMISSION_ABSORPTION) return c.free_weight ? PD_COUPLED : PD_NONE;
      return MODEL == CHI_MODEL_EMISSION_ABSORPTION_PHYSICAL ? PD_COUPLED : PD_NONE;
  }
  return PD_NONE;
}

// unconstrained u -> constrained x, with dx/du and the Jacobian term and its derivative
struct Transformed {
  double x, dxdu, logj, dlogj;
};
CHI_HD Transformed slot_backward(int dist, double u) {
  Transformed t;
  if (dist == PD_CHI2 || dist == PD_HALFNORMAL || dist == PD_TRUNCNORMAL) {
    t.x = exp(u); t.dxdu = t.x; t.logj = u; t.dlogj = 1.0;
  } else if (dist == PD_BETA || dist == PD_UNIFORM) {
    const double e = exp(-fabs(u));                 // stable sigmoid and log-sigmoids
    const double s_pos = 1.0 / (1.0 + e);           // sigmoid(|u|)
    t.x = u >= 0.0 ? s_pos : e * s_pos;
    const double l1pe = log1p(e);
    t.logj = -fabs(u) - 2.0 * l1pe;                 // log x + log(1-x)
    t.dxdu = e * s_pos * s_pos;                     // x (1 - x)
    t.dlogj = 1.0 - 2.0 * t.x;
  } else {
    t.x = u; t.dxdu = 1.0; t.logj = 0.0; t.dlogj = 0.0;
  }
  return t;
}

#define CHI_HALF_LOG_2PI 0.918938533204672741780
// log density of the free RV in slot `dist` at its constrained value x (PyMC definitions,
// third-party: restated in oracle/priors_ref.py); T is the thermal quantity the coupled
// Normal's mean depends on (tkin or fwhm2_thermal).
template <class T>
CHI_HD T slot_density(const MapCfg& c, int dist, const T& x, const T& thermal) {
  switch (dist) {
    case PD_NORMAL01: return -0.5 * (x * x) - CHI_HALF_LOG_2PI;
    case PD_CHI2: return -0.5 * m_log(x) - 0.5 * x - (0.5 * CHI_LN2 + 0.5 * 1.1447298858494001741);  // ln Gamma(1/2) = .5 ln pi
    case PD_HALFNORMAL: return -0.5 * (x * x) + 0.5 * (CHI_LN2 - 1.1447298858494001741);            // .5 ln(2/pi)
    case PD_BETA: return (c.beta_a - 1.0) * m_log(x) + (c.beta_b - 1.0) * m_log(T(1.0 - x)) - c.beta_lnB;
    case PD_UNIFORM: return 0.0 * x;
    case PD_TRUNCNORMAL: {
      T z = (x - c.nth_mu) / c.nth_sigma;
      return -0.5 * (z * z) + c.nth_const;
    }
    case PD_COUPLED: {
      T z = (x - (c.wt_mu0 - m_log10(thermal))) / c.wt_sigma;
      return -0.5 * (z * z) + c.wt_const;
    }
  }
  return 0.0 * x;
}

// physics.py:58-106 (Kim+14 eq. 4, Draine 2011 rate coefficient). The reference
// evaluates both switch branches and selects; selecting first gives the same value
// and the same (selected-branch-only) gradient.
template <class T>
CHI_HD T spin_temp(const T& tk, const T& density, const T& n_alpha) {
  T y_alpha = CHI_LYA_COEFF * n_alpha / m_powc(tk, 1.5);
  T T2 = tk / 100.0;
  T k10;
  if (m_value(tk) < 300.0) {
    k10 = 1.19e-10 * m_pow(T2, 0.74 - 0.20 * m_log(T2));
  } else {
    k10 = 2.24e-10 * m_powc(T2, 0.207) * m_exp(-0.876 / T2);
  }
  T R10 = density * k10;
  T y_c = CHI_T0 * R10 / (tk * CHI_A10);
  return (CHI_T_RADIATION + y_c * tk + y_alpha * tk) / (1.0 + y_c + y_alpha);
}

// Everything the reference registers as a per-cloud pm.Deterministic.
template <class T>
struct CloudPhys {
  T velocity, fwhm2, fwhm_L, tau_total, tspin, ff, wt;
  // extras (dead code unless chi_deterministics asks for them)
  T tkin, log10_NHI, log10_nHI, log10_depth, fwhm2_thermal, fwhm2_nonthermal;
  T log10_n_alpha, depth, TB_fwhm, thermal_fraction;
  bool has_tspin, has_ff, has_wt, has_thermal, has_TB, has_fraction;
};

// hi_physical_model.py:153-197
template <class T>
CHI_HD void physical_chain(const MapCfg& c, const T& nth_fwhm_1pc, const T& log10_n_alpha,
                           CloudPhys<T>& o) {
  o.fwhm2_nonthermal = o.fwhm2 - o.fwhm2_thermal;
  // physics.calc_depth_nonthermal (physics.py:188) on sqrt(fwhm2_nonthermal)
  T depth = m_powc(m_sqrt(o.fwhm2_nonthermal) / nth_fwhm_1pc, c.inv_power);
  o.depth = depth;
  o.log10_depth = m_log10(depth);
  o.log10_nHI = o.log10_NHI - o.log10_depth - CHI_PC_CM_LOG10;  // physics.py:225
  o.tspin = spin_temp(o.tkin, m_exp10(o.log10_nHI), m_exp10(log10_n_alpha));
  o.tau_total = m_exp10(o.log10_NHI) / o.tspin / CHI_COLUMN_CONST;  // physics.py:244
  o.has_tspin = true;
  o.has_thermal = true;
}

// Free RVs of one cloud (th[slot], constrained space) -> likelihood inputs.
// WANT_ALL also evaluates quantities the likelihood does not use.
template <int MODEL, class T, bool WANT_ALL = false>
CHI_HD CloudPhys<T> cloud_phys(const MapCfg& c, const T* th) {
  CloudPhys<T> o;
  o.has_tspin = o.has_ff = o.has_wt = o.has_thermal = o.has_TB = o.has_fraction = false;
  o.ff = T(1.0);
  o.wt = T(1.0);
  o.tspin = T(0.0);
  if constexpr (MODEL == CHI_MODEL_PHYSICS) {
    o.velocity = th[CHI_PH_VELOCITY];
    o.fwhm2 = th[CHI_PH_FWHM2];
    o.fwhm_L = th[CHI_PH_FWHM_L];
    o.tau_total = th[CHI_PH_TAU_TOTAL];
    o.tspin = th[CHI_PH_TSPIN];
    o.ff = th[CHI_PH_FILLING_FACTOR];
    o.wt = th[CHI_PH_ABSORPTION_WEIGHT];
    o.has_tspin = o.has_ff = o.has_wt = true;
    return o;
  } else {
    // hi_model.py:95-136 / hi_physical_model.py:111-151
    o.fwhm2 = c.prior_fwhm2 * th[CHI_FR_FWHM2_NORM];
    o.velocity = c.pv0 + c.pv1 * th[CHI_FR_VELOCITY_NORM];
    T log10_n_alpha = c.pa0 + c.pa1 * th[CHI_FR_LOG10_N_ALPHA_NORM];
    if constexpr (WANT_ALL) o.log10_n_alpha = log10_n_alpha;
    o.fwhm_L = c.lorentzian ? T(c.prior_fwhm_L * th[CHI_FR_FWHM_L_NORM]) : T(0.0);
    constexpr bool physical = (MODEL >= CHI_MODEL_ABSORPTION_PHYSICAL);
    if constexpr (!physical) o.log10_nHI = c.pn0 + c.pn1 * th[CHI_FR_LOG10_NHI_NORM];

    if constexpr (MODEL == CHI_MODEL_ABSORPTION) {
      // absorption_model.py:44-72; tspin is a Deterministic only, not in the likelihood
      o.tau_total = c.prior_amp * th[CHI_FR_AMPLITUDE_NORM];
      if constexpr (WANT_ALL) {
        o.tkin = th[CHI_FR_THERMAL] * (o.fwhm2 / (CHI_THERMAL_CONST * CHI_THERMAL_CONST));
        o.tspin = spin_temp(o.tkin, m_exp10(o.log10_nHI), m_exp10(log10_n_alpha));
        o.has_tspin = true;
      }
    } else if constexpr (MODEL == CHI_MODEL_EMISSION || MODEL == CHI_MODEL_EMISSION_ABSORPTION) {
      // emission_model.py:60-132
      T TB_fwhm = c.prior_amp * th[CHI_FR_AMPLITUDE_NORM];
      if constexpr (WANT_ALL) {
        o.TB_fwhm = TB_fwhm;
        o.has_TB = true;
      }
      T root = m_sqrt(o.fwhm2);
      T tkin_min = TB_fwhm / root;
      T tkin_max = o.fwhm2 / (CHI_THERMAL_CONST * CHI_THERMAL_CONST);
      o.tkin = m_select(m_value(tkin_min) > m_value(tkin_max), tkin_max,
                        T(th[CHI_FR_THERMAL] * (tkin_max - tkin_min) + tkin_min));
      o.tspin = spin_temp(o.tkin, m_exp10(o.log10_nHI), m_exp10(log10_n_alpha));
      T ff_min = tkin_min / o.tspin;
      o.ff = m_select(m_value(ff_min) > 1.0, T(1.0),
                      T(th[CHI_FR_FILLING_FACTOR_NORM] * (1.0 - ff_min) + ff_min));
      T x = m_clip(T(ff_min / o.ff), 0.0, 0.9999);
      T tau_peak = -m_log(T(1.0 - x));
      o.tau_total = tau_peak * root * c.const_G;
      o.has_tspin = o.has_ff = true;
      if constexpr (MODEL == CHI_MODEL_EMISSION_ABSORPTION) {
        // emission_absorption_model.py:45-64
        if (c.free_weight) o.wt = m_exp10(th[CHI_FR_LOG10_WT]) * o.ff * o.tkin;
        o.has_wt = true;
      }
    } else if constexpr (MODEL == CHI_MODEL_ABSORPTION_PHYSICAL) {
      // absorption_physical_model.py:43-77
      T NHI_fwhm2_thermal = c.prior_amp * th[CHI_FR_AMPLITUDE_NORM];
      o.fwhm2_thermal = o.fwhm2 * th[CHI_FR_THERMAL];
      o.tkin = o.fwhm2_thermal / (CHI_THERMAL_CONST * CHI_THERMAL_CONST);
      o.log10_NHI = m_log10(T(NHI_fwhm2_thermal * o.fwhm2_thermal));
      physical_chain(c, th[CHI_FR_NTH_FWHM_1PC], log10_n_alpha, o);
    } else {
      // emission_physical_model.py:59-129
      T ff_NHI = c.prior_amp * th[CHI_FR_AMPLITUDE_NORM];
      T tkin_min = (ff_NHI / CHI_COLUMN_CONST) / (c.const_G * m_sqrt(o.fwhm2));
      T frac_min = (CHI_THERMAL_CONST * CHI_THERMAL_CONST) * tkin_min / o.fwhm2;
      T frac = m_select(m_value(frac_min) > 1.0, T(1.0),
                        T(th[CHI_FR_THERMAL] * (1.0 - frac_min) + frac_min));
      if constexpr (WANT_ALL) {
        o.thermal_fraction = frac;
        o.has_fraction = true;
      }
      o.fwhm2_thermal = o.fwhm2 * frac;
      o.tkin = o.fwhm2_thermal / (CHI_THERMAL_CONST * CHI_THERMAL_CONST);
      T ff_min = tkin_min / o.tkin;
      o.ff = m_select(m_value(ff_min) > 1.0, T(1.0),
                      T(th[CHI_FR_FILLING_FACTOR_NORM] * (1.0 - ff_min) + ff_min));
      o.log10_NHI = m_log10(T(ff_NHI / o.ff));
      physical_chain(c, th[CHI_FR_NTH_FWHM_1PC], log10_n_alpha, o);
      o.has_ff = true;
      if constexpr (MODEL == CHI_MODEL_EMISSION_ABSORPTION_PHYSICAL) {
        // emission_absorption_physical_model.py:41-60
        o.wt = m_exp10(th[CHI_FR_LOG10_WT]) * o.ff * o.fwhm2_thermal;
        o.has_wt = true;
      }
    }
    if constexpr (!physical && WANT_ALL) {
      // hi_model.py:138-162
      if (o.has_tspin) {
        o.log10_NHI = m_log10(T(CHI_COLUMN_CONST * o.tspin * o.tau_total));  // physics.py:207
        o.log10_depth = o.log10_NHI - o.log10_nHI - CHI_PC_CM_LOG10;
      }
    }
    return o;
  }
}

// Line-profile constants of one cloud on one dataset (physics.py:294-309):
//   tau(v) = AG * exp(-k (v-c)^2) + AL / ((v-c)^2 + h)
template <class T>
struct ProfConst {
  T k, AG, AL, h;
};

template <class T>
CHI_HD ProfConst<T> profile_consts(const T& fwhm2, const T& fwhm_L, const T& amp, double wch2,
                                   bool lorentzian) {
  ProfConst<T> p;
  T W2 = fwhm2 + wch2 + fwhm_L * fwhm_L;  // fwhm_conv**2, physics.py:296
  T W = m_sqrt(W2);
  p.k = (4.0 * CHI_LN2) / W2;                            // physics.py:33
  T g0 = m_sqrt(T((4.0 * CHI_LN2 / CHI_PI) / W2));       // physics.py:33-35
  p.h = 0.25 * W2;                                       // physics.py:55
  if (lorentzian) {
    T r = fwhm_L / W;                                    // physics.py:297
    T eta = 1.36603 * r - 0.47719 * (r * r) + 0.11116 * (r * r * r);  // physics.py:298-300
    p.AG = amp * (1.0 - eta) * g0;
    p.AL = amp * eta * W / (2.0 * CHI_PI);
  } else {
    p.AG = amp * g0;
    p.AL = T(0.0);
  }
  return p;
}

// Packed per-(draw, cloud) constants handed from the map kernel to the channel kernels.
#define CHI_CC_STRIDE 12
enum { CC_C = 0, CC_TS, CC_F, CC_KE, CC_AGE, CC_ALE, CC_HE, CC_KA, CC_AGA, CC_ALA, CC_HA, CC_CAREFUL };

}  // namespace chi
