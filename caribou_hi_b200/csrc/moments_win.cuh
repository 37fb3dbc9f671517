// Windowed variants of the two hot kernels (same outputs and record layout as moments.cuh).
//
// Two changes against k_moments_em / k_moments_abs:
//  * CPT channels per thread and iteration give the instruction-level parallelism (independent
//    exp chains) instead of the unrolled cloud loop;
//  * for the pure Gaussian profile each cloud gets an *active channel window*: outside it the
//    cloud's optical depth is below 2^-54 (so 1 - exp(-tau) is exactly 0 in fp64, the transfer is
//    exactly the identity) and exp(-k d^2) < 2^-70 (its gradient terms are < 1e-15 of the kept
//    ones), and the cloud is skipped for that warp iteration with a warp-uniform branch.  Line
//    wings are most of a wide spectrum, so this removes a data-dependent share of the work
//    without changing results beyond fp64 rounding.  Windows are found once per warp by
//    binary search on the (monotonic) spectral axis; any non-finite constant disables the
//    window for that cloud so that NaNs propagate exactly as in the reference.
#pragma once
#include "moments.cuh"

namespace chi {

constexpr int em2_min_blocks(int N, bool pv, int cpt) {
  const int regs = 2 * ((pv ? 8 : 5) * N + (pv ? 4 : 3) * N * cpt + 8 * cpt) + 48;
  return clamp_blocks((regs <= 128 ? 4 : (regs <= 168 ? 3 : 2)) + CHI_BLOCKS_BIAS);
}
constexpr int ab2_min_blocks(int N, bool pv, int cpt) {
  const int regs = 2 * ((pv ? 6 : 3) * N + (pv ? 3 : 2) * N * cpt + 6 * cpt) + 48;
  return clamp_blocks((regs <= 128 ? 4 : (regs <= 168 ? 3 : 2)) + CHI_BLOCKS_BIAS);
}

// Active window [lo, hi] (channel indices, inclusive) of a Gaussian cloud on a monotonic axis:
// k d^2 <= L with L = max(48.5, 37.43 + ln|AG|)  (2^-70 and 2^-54/|AG|), widened by one channel.
__device__ __forceinline__ void cloud_window(const double* __restrict__ V, int S, int dir, double c, double k,
                                            double AG, bool finite_rest, int& lo, int& hi) {
  lo = 0;
  hi = S - 1;
  if (dir == 0) return;
  const double L = fmax(48.5, 37.43 + log(fabs(AG)));
  const double R = sqrt(L / k);
  if (!(finite_rest && isfinite(R) && isfinite(c) && isfinite(AG))) return;
  const double key_c = dir * c;
  auto lower = [&](double x) {  // first index with dir*V[i] >= x
    int l = 0, h = S;
    while (l < h) {
      const int m = (l + h) >> 1;
      if (dir * V[m] < x) l = m + 1; else h = m;
    }
    return l;
  };
  lo = max(0, lower(key_c - R) - 1);
  hi = min(S - 1, lower(key_c + R));
}

// ---------------------------------------------------------------------------
// emission
// ---------------------------------------------------------------------------
template <int N, bool PV, int CPT>
__global__ void __launch_bounds__(CHI_BLOCK, em2_min_blocks(N, PV, CPT)) k_moments_em2(ChanArgs a) {
  constexpr int NE = EmLayout<PV>::NE;
  __shared__ __align__(16) double s_c[CHI_WARPS_PER_BLOCK][N][8];
  __shared__ int2 s_win[CHI_WARPS_PER_BLOCK][N];
  __shared__ double s_gb[CHI_MAX_BASELINE - CHI_BASE_REG][CHI_BLOCK];
  __shared__ double s_beta[CHI_WARPS_PER_BLOCK][CHI_MAX_BASELINE];
  __shared__ double s_T[64];
  fill_exp_table(s_T);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t item = (int64_t)blockIdx.x * CHI_WARPS_PER_BLOCK + warp;
  if (item >= a.B * a.n_chunks) return;
  const int64_t b = item / a.n_chunks;
  const int chunk = (int)(item - b * a.n_chunks);
  const int nb = a.n_base, S = a.S;
  const double* __restrict__ V = a.vax;

  if (lane < N) {
    const double* p = a.cc + (b * N + lane) * CHI_CC_STRIDE;
    const double f = p[CC_F], ts = p[CC_TS], c0 = p[CC_C], k = p[CC_KE], AG = p[CC_AGE];
    double* d = s_c[warp][lane];
    d[0] = c0; d[1] = -k; d[2] = -AG; d[3] = -p[CC_ALE]; d[4] = p[CC_HE];
    d[5] = f; d[6] = ts; d[7] = f * ts;
    int lo, hi;
    cloud_window(V, S, PV ? 0 : a.axis_dir, c0, k, AG, isfinite(f) && isfinite(ts), lo, hi);
    s_win[warp][lane] = make_int2(lo, hi);
  }
  if (lane < nb) s_beta[warp][lane] = a.coeff ? a.coeff[b * nb + lane] : 0.0;
  for (int i = CHI_BASE_REG; i < nb; ++i) s_gb[i - CHI_BASE_REG][threadIdx.x] = 0.0;
  __syncwarp();
  const double beta0 = nb > 0 ? s_beta[warp][0] : 0.0, beta1 = nb > 1 ? s_beta[warp][1] : 0.0;
  const double* __restrict__ bas0 = a.basis;
  const double* __restrict__ bas1 = a.basis + S;

  const int sl = a.sl ? a.sl[b] : 0;
  const double* __restrict__ Y = a.yeff + (int64_t)sl * S;
  const double* __restrict__ Wt = a.wgt + (int64_t)sl * S;
  const double* __restrict__ G = a.gbar ? a.gbar + b * S : nullptr;

  double acc[N][NE];
#pragma unroll
  for (int n = 0; n < N; ++n)
#pragma unroll
    for (int j = 0; j < NE; ++j) acc[n][j] = 0.0;
  double lp = 0.0, gb0 = 0.0, gb1 = 0.0;
  const double bg = a.bg;
  const double(*sc)[8] = s_c[warp];
  const ExpCtx ek = CHI_LOAD_EXPCTX(s_T);

  const int s_end = min(S, (chunk + 1) * a.chunk_len);
  for (int s0 = chunk * a.chunk_len; s0 < s_end; s0 += 32 * CPT) {
    const int w1 = min(s0 + 32 * CPT, s_end) - 1;  // last channel of this warp iteration
    int si[CPT];
    bool ok[CPT];
    double v[CPT], base[CPT];
#pragma unroll
    for (int c = 0; c < CPT; ++c) {
      const int s = s0 + lane + 32 * c;
      ok[c] = s < s_end;
      si[c] = ok[c] ? s : s_end - 1;
      v[c] = V[si[c]];
      base[c] = 0.0;
      if (nb > 0) base[c] = beta0 * bas0[si[c]];
      if (nb > 1) base[c] = fma(beta1, bas1[si[c]], base[c]);
      for (int i = CHI_BASE_REG; i < nb; ++i)
        base[c] = fma(s_beta[warp][i], a.basis[(int64_t)i * S + si[c]], base[c]);
    }

    // ---- forward, nearest cloud first ----
    double om[N][CPT], Pn[N][CPT], E[N][CPT], q[N][CPT];
    double P[CPT], T[CPT];
#pragma unroll
    for (int c = 0; c < CPT; ++c) { P[c] = 1.0; T[c] = 0.0; }
    unsigned act = 0;
#pragma unroll
    for (int n = 0; n < N; ++n) {
      const int2 w = s_win[warp][n];
      if (PV || (s0 <= w.y && w1 >= w.x)) {
        act |= 1u << n;
        double c0, nk, nAG, nAL, h, f, Ts, fTs;
        lds2(&sc[n][0], c0, nk);
        lds2(&sc[n][2], nAG, nAL);
        lds2(&sc[n][4], h, f);
        lds2(&sc[n][6], Ts, fTs);
#pragma unroll
        for (int c = 0; c < CPT; ++c) {
          const double d = v[c] - c0, d2 = d * d;
          E[n][c] = chi_exp<false>(nk * d2, ek);
          double x = nAG * E[n][c];             // -tau
          if (PV) {
            q[n][c] = chi_rcp(d2 + h);
            x = fma(nAL, q[n][c], x);
          }
          om[n][c] = 1.0 - chi_exp<true>(x, ek);  // 1 - e^-tau
          Pn[n][c] = P[c];
          T[c] = fma(fTs * om[n][c], P[c], T[c]);
          P[c] *= fma(-f, om[n][c], 1.0);
        }
      }
    }

    double mubar[CPT];
#pragma unroll
    for (int c = 0; c < CPT; ++c) {
      const double mu = (bg * P[c] + T[c]) - bg + base[c];  // ON - OFF, physics.py:362-365
      if (G) {
        mubar[c] = ok[c] ? G[si[c]] : 0.0;
      } else {
        const double r = Y[si[c]] - mu;
        mubar[c] = ok[c] ? Wt[si[c]] * r : 0.0;
        lp = fma(-0.5 * mubar[c], r, lp);
      }
      if (nb > 0) gb0 = fma(mubar[c], bas0[si[c]], gb0);
      if (nb > 1) gb1 = fma(mubar[c], bas1[si[c]], gb1);
      for (int i = CHI_BASE_REG; i < nb; ++i)
        s_gb[i - CHI_BASE_REG][threadIdx.x] =
            fma(mubar[c], a.basis[(int64_t)i * S + si[c]], s_gb[i - CHI_BASE_REG][threadIdx.x]);
    }

    // ---- reverse, farthest cloud first; Q = brightness arriving behind cloud n ----
    double Q[CPT];
#pragma unroll
    for (int c = 0; c < CPT; ++c) Q[c] = bg;
#pragma unroll
    for (int n = N - 1; n >= 0; --n) {
      if (PV || ((act >> n) & 1u)) {
        double c0, nk, h, f, Ts, fTs;
        lds2(&sc[n][0], c0, nk);
        lds2(&sc[n][4], h, f);
        lds2(&sc[n][6], Ts, fTs);
#pragma unroll
        for (int c = 0; c < CPT; ++c) {
          const double Sn = Ts - Q[c];
          const double x = om[n][c] * Pn[n][c];
          const double mx = mubar[c] * x;
          acc[n][3] += mx;
          acc[n][4] = fma(mx, Sn, acc[n][4]);
          const double g = (mubar[c] * Sn) * (Pn[n][c] - x);  // taubar / f
          const double d = v[c] - c0;
          const double t = g * E[n][c];
          const double td = t * d;
          acc[n][0] += t;
          acc[n][1] += td;
          acc[n][2] = fma(td, d, acc[n][2]);
          if (PV) {
            const double u = g * q[n][c];
            const double u2 = u * q[n][c];
            acc[n][5] += u;
            acc[n][6] += u2;
            acc[n][7] = fma(u2, d, acc[n][7]);
          }
          Q[c] = fma(fma(-f, om[n][c], 1.0), Q[c], fTs * om[n][c]);
        }
      }
    }
  }

  double* out = a.mom + item * a.mom_stride;
  lp = warp_sum(lp);
  gb0 = warp_sum(gb0);
  gb1 = warp_sum(gb1);
  if (lane == 0) {
    out[0] = lp;
    if (nb > 0) out[1] = gb0;
    if (nb > 1) out[2] = gb1;
  }
  for (int i = CHI_BASE_REG; i < nb; ++i) {
    double g = warp_sum(s_gb[i - CHI_BASE_REG][threadIdx.x]);
    if (lane == 0) out[1 + i] = g;
  }
  double* om_ = out + 1 + nb;
#pragma unroll
  for (int n = 0; n < N; ++n)
#pragma unroll
    for (int j = 0; j < NE; ++j) {
      double r = warp_sum(acc[n][j]);
      if (lane == ((n * NE + j) & 31)) om_[n * NE + j] = r;
    }
}

// ---------------------------------------------------------------------------
// absorption
// ---------------------------------------------------------------------------
template <int N, bool PV, int CPT>
__global__ void __launch_bounds__(CHI_BLOCK, ab2_min_blocks(N, PV, CPT)) k_moments_abs2(ChanArgs a) {
  constexpr int NA = AbLayout<PV>::NA;
  __shared__ __align__(16) double s_c[CHI_WARPS_PER_BLOCK][N][6];
  __shared__ int2 s_win[CHI_WARPS_PER_BLOCK][N];
  __shared__ double s_gb[CHI_MAX_BASELINE - CHI_BASE_REG][CHI_BLOCK];
  __shared__ double s_beta[CHI_WARPS_PER_BLOCK][CHI_MAX_BASELINE];
  __shared__ double s_T[64];
  fill_exp_table(s_T);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t item = (int64_t)blockIdx.x * CHI_WARPS_PER_BLOCK + warp;
  if (item >= a.B * a.n_chunks) return;
  const int64_t b = item / a.n_chunks;
  const int chunk = (int)(item - b * a.n_chunks);
  const int nb = a.n_base, S = a.S;
  const double* __restrict__ V = a.vax;

  if (lane < N) {
    const double* p = a.cc + (b * N + lane) * CHI_CC_STRIDE;
    const double c0 = p[CC_C], k = p[CC_KA], AG = p[CC_AGA];
    double* d = s_c[warp][lane];
    d[0] = c0; d[1] = -k; d[2] = -AG; d[3] = -p[CC_ALA]; d[4] = p[CC_HA]; d[5] = 0.0;
    int lo, hi;
    cloud_window(V, S, PV ? 0 : a.axis_dir, c0, k, AG, true, lo, hi);
    s_win[warp][lane] = make_int2(lo, hi);
  }
  if (lane < nb) s_beta[warp][lane] = a.coeff ? a.coeff[b * nb + lane] : 0.0;
  for (int i = CHI_BASE_REG; i < nb; ++i) s_gb[i - CHI_BASE_REG][threadIdx.x] = 0.0;
  __syncwarp();
  const double beta0 = nb > 0 ? s_beta[warp][0] : 0.0, beta1 = nb > 1 ? s_beta[warp][1] : 0.0;
  const double* __restrict__ bas0 = a.basis;
  const double* __restrict__ bas1 = a.basis + S;

  const int sl = a.sl ? a.sl[b] : 0;
  const double* __restrict__ Y = a.yeff + (int64_t)sl * S;
  const double* __restrict__ Wt = a.wgt + (int64_t)sl * S;
  const double* __restrict__ G = a.gbar ? a.gbar + b * S : nullptr;

  double acc[N][NA];
#pragma unroll
  for (int n = 0; n < N; ++n)
#pragma unroll
    for (int j = 0; j < NA; ++j) acc[n][j] = 0.0;
  double lp = 0.0, gb0 = 0.0, gb1 = 0.0;
  const double(*sc)[6] = s_c[warp];
  const ExpCtx ek = CHI_LOAD_EXPCTX(s_T);

  const int s_end = min(S, (chunk + 1) * a.chunk_len);
  for (int s0 = chunk * a.chunk_len; s0 < s_end; s0 += 32 * CPT) {
    const int w1 = min(s0 + 32 * CPT, s_end) - 1;
    int si[CPT];
    bool ok[CPT];
    double v[CPT], base[CPT], xs[CPT];
#pragma unroll
    for (int c = 0; c < CPT; ++c) {
      const int s = s0 + lane + 32 * c;
      ok[c] = s < s_end;
      si[c] = ok[c] ? s : s_end - 1;
      v[c] = V[si[c]];
      xs[c] = 0.0;  // -sum_n tau_n
      base[c] = 0.0;
      if (nb > 0) base[c] = beta0 * bas0[si[c]];
      if (nb > 1) base[c] = fma(beta1, bas1[si[c]], base[c]);
      for (int i = CHI_BASE_REG; i < nb; ++i)
        base[c] = fma(s_beta[warp][i], a.basis[(int64_t)i * S + si[c]], base[c]);
    }

    double E[N][CPT], q[N][CPT];
    unsigned act = 0;
#pragma unroll
    for (int n = 0; n < N; ++n) {
      const int2 w = s_win[warp][n];
      if (PV || (s0 <= w.y && w1 >= w.x)) {
        act |= 1u << n;
        double c0, nk, nAG, nAL;
        lds2(&sc[n][0], c0, nk);
        lds2(&sc[n][2], nAG, nAL);
#pragma unroll
        for (int c = 0; c < CPT; ++c) {
          const double d = v[c] - c0, d2 = d * d;
          E[n][c] = chi_exp<false>(nk * d2, ek);
          xs[c] = fma(nAG, E[n][c], xs[c]);
          if (PV) {
            q[n][c] = chi_rcp(d2 + sc[n][4]);
            xs[c] = fma(nAL, q[n][c], xs[c]);
          }
        }
      }
    }
    double g[CPT];
#pragma unroll
    for (int c = 0; c < CPT; ++c) {
      const double ea = chi_exp<true>(xs[c], ek);
      const double mu = (1.0 - ea) + base[c];
      double mubar;
      if (G) {
        mubar = ok[c] ? G[si[c]] : 0.0;
      } else {
        const double r = Y[si[c]] - mu;
        mubar = ok[c] ? Wt[si[c]] * r : 0.0;
        lp = fma(-0.5 * mubar, r, lp);
      }
      if (nb > 0) gb0 = fma(mubar, bas0[si[c]], gb0);
      if (nb > 1) gb1 = fma(mubar, bas1[si[c]], gb1);
      for (int i = CHI_BASE_REG; i < nb; ++i)
        s_gb[i - CHI_BASE_REG][threadIdx.x] =
            fma(mubar, a.basis[(int64_t)i * S + si[c]], s_gb[i - CHI_BASE_REG][threadIdx.x]);
      g[c] = mubar * ea;  // taubar, identical for every cloud
    }
#pragma unroll
    for (int n = 0; n < N; ++n) {
      if (PV || ((act >> n) & 1u)) {
        const double c0 = sc[n][0];
#pragma unroll
        for (int c = 0; c < CPT; ++c) {
          const double d = v[c] - c0;
          const double t = g[c] * E[n][c];
          const double td = t * d;
          acc[n][0] += t;
          acc[n][1] += td;
          acc[n][2] = fma(td, d, acc[n][2]);
          if (PV) {
            const double u = g[c] * q[n][c];
            const double u2 = u * q[n][c];
            acc[n][3] += u;
            acc[n][4] += u2;
            acc[n][5] = fma(u2, d, acc[n][5]);
          }
        }
      }
    }
  }

  double* out = a.mom + item * a.mom_stride;
  lp = warp_sum(lp);
  gb0 = warp_sum(gb0);
  gb1 = warp_sum(gb1);
  if (lane == 0) {
    out[0] = lp;
    if (nb > 0) out[1] = gb0;
    if (nb > 1) out[2] = gb1;
  }
  for (int i = CHI_BASE_REG; i < nb; ++i) {
    double gsum = warp_sum(s_gb[i - CHI_BASE_REG][threadIdx.x]);
    if (lane == 0) out[1 + i] = gsum;
  }
  double* om_ = out + 1 + nb;
#pragma unroll
  for (int n = 0; n < N; ++n)
#pragma unroll
    for (int j = 0; j < NA; ++j) {
      double r = warp_sum(acc[n][j]);
      if (lane == ((n * NA + j) & 31)) om_[n * NA + j] = r;
    }
}

}  // namespace chi
