/*
 * chi_oracle.c -- plain-C CPU restatement of the likelihood hot path (TEST INFRASTRUCTURE and
 * CPU baseline; never linked into the product).
 *
 * Forward arithmetic follows the reference line by line:
 *   caribou_hi/physics.py:294-309  calc_pseudo_voigt  (channel width, convolved FWHM, eta, G, L)
 *   caribou_hi/physics.py:345-365  radiative_transfer (attenuation cumprod, ON - OFF)
 *   emission_absorption_model.py:83-99 (tau_e = tau_total*phi_e, tau_a = wt*tau_total*phi_a,
 *   1 - exp(-sum tau_a)) and pm.Normal logp (PyMC, third-party).
 * The gradient is the hand-derived reverse pass of SURVEY.md Appendix B written as straight
 * loops over [S][N] temporaries, the way PyTensor's fused C loops evaluate the reference graph
 * (one evaluation per call, no SIMD intrinsics); OpenMP over independent draws stands in for
 * PyMC's process-per-chain parallelism.  It is cross-checked against the torch-autograd
 * oracle in tests/test_c_oracle.py.
 *
 * Inputs per draw: the seven likelihood inputs in[7][N] (velocity, fwhm2, fwhm_L, tau_total,
 * tspin, filling_factor, absorption_weight); outputs logp and grad[7][N].
 * Build: gcc -O3 -fopenmp -shared -fPIC oracle/chi_oracle.c -o oracle/_build/libchi_oracle.so -lm
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define MAXN 16

typedef struct {
  int S;              /* channels (0 = dataset absent) */
  const double* v;    /* [S] spectral axis */
  const double* y;    /* [S] brightness minus the constant baseline */
  const double* sig;  /* [S] noise */
} oracle_dataset;

static const double LN2 = 0.693147180559945309417;
static const double PI = 3.14159265358979323846;

/* profile and its partials for one cloud on one axis; returns fwhm_conv */
static void profile_consts(double dv, double fwhm2, double fwhm_L, double* W, double* eta) {
  const double channel_fwhm = 4.0 * LN2 * fabs(dv) / PI;                      /* physics.py:294-295 */
  *W = sqrt(fwhm2 + channel_fwhm * channel_fwhm + fwhm_L * fwhm_L);           /* :296 */
  const double r = fwhm_L / *W;                                               /* :297 */
  *eta = 1.36603 * r - 0.47719 * r * r + 0.11116 * r * r * r;                 /* :298-300 */
}

/* one draw: logp and gradient w.r.t. in[7][N] */
static void one_draw(int N, const oracle_dataset* e, const oracle_dataset* a, double bg, const double* in,
                     double* logp_out, double* grad) {
  const double *velocity = in, *fwhm2 = in + N, *fwhm_L = in + 2 * N, *tau_total = in + 3 * N,
               *tspin = in + 4 * N, *ff = in + 5 * N, *wt = in + 6 * N;
  double g_c[MAXN] = {0}, g_W[MAXN] = {0}, g_eta[MAXN] = {0}, g_amp_e[MAXN] = {0}, g_amp_a[MAXN] = {0},
         g_ts[MAXN] = {0}, g_ff[MAXN] = {0};
  double W_[2][MAXN], eta_[2][MAXN];
  double logp = 0.0;
  const double half_log_2pi = 0.5 * log(2.0 * PI);

  for (int ds = 0; ds < 2; ++ds) {
    const oracle_dataset* d = ds == 0 ? e : a;
    if (!d || d->S == 0) continue;
    const double dv = d->v[1] - d->v[0];
    for (int n = 0; n < N; ++n) profile_consts(dv, fwhm2[n], fwhm_L[n], &W_[ds][n], &eta_[ds][n]);
    for (int s = 0; s < d->S; ++s) {
      double G[MAXN], L[MAXN], tau[MAXN], ex[MAXN], att[MAXN + 1], dd[MAXN];
      double mu;
      /* forward */
      for (int n = 0; n < N; ++n) {
        const double W = W_[ds][n], x = d->v[s] - velocity[n];
        dd[n] = x;
        G[n] = exp(-4.0 * LN2 * x * x / (W * W)) * sqrt(4.0 * LN2 / (PI * W * W));   /* physics.py:33-35 */
        L[n] = W / (2.0 * PI) / (x * x + (W / 2.0) * (W / 2.0));                     /* :55 */
        const double phi = eta_[ds][n] * L[n] + (1.0 - eta_[ds][n]) * G[n];          /* :309 */
        const double amp = ds == 0 ? tau_total[n] : wt[n] * tau_total[n];
        tau[n] = amp * phi;
      }
      if (ds == 0) {
        att[0] = 1.0;
        double sum = 0.0;
        for (int n = 0; n < N; ++n) {
          ex[n] = exp(-tau[n]);
          att[n + 1] = att[n] * (1.0 - ff[n] + ff[n] * ex[n]);                       /* physics.py:345-351 */
          sum += ff[n] * tspin[n] * (1.0 - ex[n]) * att[n];                          /* :357-361 */
        }
        mu = bg * att[N] + sum - bg;                                                /* :362-365 */
      } else {
        double tsum = 0.0;
        for (int n = 0; n < N; ++n) tsum += tau[n];
        ex[0] = exp(-tsum);
        mu = 1.0 - ex[0];                                                           /* absorption_model.py:91 */
      }
      const double z = (d->y[s] - mu) / d->sig[s];
      logp += -0.5 * z * z - half_log_2pi - log(d->sig[s]);                         /* pm.Normal logp */
      const double mubar = (d->y[s] - mu) / (d->sig[s] * d->sig[s]);
      /* reverse */
      double Q = bg;
      for (int n = N - 1; n >= 0; --n) {
        double taubar;
        if (ds == 0) {
          const double one_m_e = 1.0 - ex[n];
          g_ts[n] += mubar * ff[n] * one_m_e * att[n];
          g_ff[n] += mubar * one_m_e * att[n] * (tspin[n] - Q);
          taubar = mubar * ex[n] * att[n] * ff[n] * (tspin[n] - Q);
          Q = ff[n] * tspin[n] * one_m_e + (1.0 - ff[n] + ff[n] * ex[n]) * Q;
        } else {
          taubar = mubar * ex[0];
        }
        const double W = W_[ds][n], eta = eta_[ds][n], x = dd[n];
        const double amp = ds == 0 ? tau_total[n] : wt[n] * tau_total[n];
        const double phi = eta * L[n] + (1.0 - eta) * G[n];
        if (ds == 0) g_amp_e[n] += taubar * phi; else g_amp_a[n] += taubar * phi;
        const double k = 4.0 * LN2 / (W * W), q = 1.0 / (x * x + W * W / 4.0);
        const double phibar = taubar * amp;
        /* d phi / d center, d phi / d W (eta fixed), d phi / d eta */
        g_c[n] += phibar * ((1.0 - eta) * G[n] * 2.0 * k * x + eta * L[n] * q * 2.0 * x);
        g_W[n] += phibar * ((1.0 - eta) * G[n] * (2.0 * k * x * x / W - 1.0 / W) +
                            eta * (L[n] / W - L[n] * q * W / 2.0));
        g_eta[n] += phibar * (L[n] - G[n]);
      }
    }
    /* chain rule W, eta -> fwhm2, fwhm_L for this dataset (physics.py:296-300) */
    for (int n = 0; n < N; ++n) {
      const double W = W_[ds][n], r = fwhm_L[n] / W;
      const double rbar = g_eta[n] * (1.36603 - 2.0 * 0.47719 * r + 3.0 * 0.11116 * r * r);
      const double Wbar = g_W[n] + rbar * (-fwhm_L[n] / (W * W));
      grad[1 * N + n] += Wbar / (2.0 * W);
      grad[2 * N + n] += rbar / W + Wbar * fwhm_L[n] / W;
      g_W[n] = 0.0;
      g_eta[n] = 0.0;
    }
  }
  for (int n = 0; n < N; ++n) {
    grad[0 * N + n] += g_c[n];
    grad[3 * N + n] += g_amp_e[n] + wt[n] * g_amp_a[n];
    grad[4 * N + n] += g_ts[n];
    grad[5 * N + n] += g_ff[n];
    grad[6 * N + n] += tau_total[n] * g_amp_a[n];
  }
  *logp_out = logp;
}

/* B draws; in[B][7][N], logp[B], grad[B][7][N].  Returns the number of threads used. */
int chi_oracle_logp_grad(int B, int N, const oracle_dataset* emission, const oracle_dataset* absorption,
                         double bg_temp, const double* in, double* logp, double* grad, int n_threads) {
  if (N > MAXN) return -1;
  int used = 1;
#ifdef _OPENMP
  if (n_threads > 0) omp_set_num_threads(n_threads);
  used = omp_get_max_threads();
#endif
  memset(grad, 0, (size_t)B * 7 * N * sizeof(double));
#pragma omp parallel for schedule(static)
  for (int b = 0; b < B; ++b)
    one_draw(N, emission, absorption, bg_temp, in + (size_t)b * 7 * N, logp + b, grad + (size_t)b * 7 * N);
  return used;
}
