// The two O(B N S) kernels of the hot path (sm_100a, FP64 pipe bound):
//   k_moments_em   emission spectrum:  profile -> tau -> ordered-cloud radiative transfer
//                  -> residual -> reverse pass, fused in registers per channel
//   k_moments_abs  absorption spectrum: 1 - exp(-sum_n tau_n)
// One warp owns one (draw, channel chunk); lanes walk channels s = lane, lane+32, ...
// (coalesced 8-byte loads of the axis, spectrum and weights), and accumulate per-cloud
// channel sums ("moments") which k_epilogue turns into parameter gradients.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "cloud_map.cuh"

namespace chi {

#ifndef CHI_WARPS_PER_BLOCK
#define CHI_WARPS_PER_BLOCK 4
#endif
#define CHI_BLOCK (32 * CHI_WARPS_PER_BLOCK)
#define CHI_FULL 0xffffffffu

// exp() of the channel path.  The arguments are -k d^2 and -tau, i.e. <= 0 on every
// model-level path, and a large share of them sits far below the underflow threshold
// (line wings), so the libm slow path (a divergent branch) is replaced by a select:
//   * n = rint(x log2 e) by the 1.5*2^52 trick, r = x - n ln2 (hi/lo split),
//   * exp(r) = 1 + r + r^2 g(r), g a degree-9 polynomial fitted to (e^r-1-r)/r^2 on
//     |r| <= ln2/2 (max relative error 1.6e-17, below fp64 rounding),
//   * 2^n applied by adding n to the exponent field,
//   * x < -708 (result would be subnormal, < 2.3e-308) returns 0; x >= 700, +inf and
//     NaN with the sign bit clear take libm's exp (never happens for tau >= 0).
// coefficients in the constant bank: FP64 literals would be re-materialised by two UMOVs at
// every use once the register cap stops the compiler from keeping them resident
__constant__ double c_exp[16] = {
    1.4426950408889634074,       // 0 log2(e)
    6755399441055744.0,          // 1 1.5 * 2^52
    -6.93147180369123816490e-01, // 2 -ln2 hi
    -1.90821492927058770002e-10, // 3 -ln2 lo
    2.510038549551032e-08,       // 4.. g(r) high -> low degree
    2.7620088445409746e-07, 2.7557268459997064e-06, 2.4801521295954376e-05, 0.00019841269863053618,
    0.0013888888917213717, 0.008333333333330062, 0.04166666666662413, 0.16666666666666669,
    0.5000000000000001, 0.0, 0.0};

// The coefficients, read once per kernel into (uniform) registers.
struct ExpK {
  double k[14];
};
__device__ __forceinline__ ExpK load_expk() {
  ExpK e;
#pragma unroll
  for (int i = 0; i < 14; ++i) e.k[i] = c_exp[i];
  return e;
}

// Branch-free on purpose: a branch per call would put every exp in its own basic block and
// stop the scheduler from interleaving the independent per-cloud chains (ILP).
// OVERFLOW=false is for arguments that are <= 0 by construction (-k d^2).
template <bool OVERFLOW = true>
__device__ __forceinline__ double chi_exp(double x, const ExpK& c) {
  const int hi = __double2hiint(x);
  const double t = fma(x, c.k[0], c.k[1]);
  const double nf = t - c.k[1];
  double r = fma(nf, c.k[2], x);
  r = fma(nf, c.k[3], r);
  double p = fma(c.k[4], r, c.k[5]);
#pragma unroll
  for (int i = 6; i <= 13; ++i) p = fma(p, r, c.k[i]);
  p = fma(p, r, 1.0);
  p = fma(p, r, 1.0);
  const int n = __double2loint(t);
  const double res = __hiloint2double(__double2hiint(p) + (n << 20), __double2loint(p));
  // -708 = 0xC086200000000000; negative doubles order like unsigned ints on the high word;
  // the range stops at -inf (0xFFF00000) so that NaNs still propagate through r and p
  const bool underflow = (unsigned)(hi - 0xC0862001) <= (0xFFF00000u - 0xC0862001u);
  double out = underflow ? 0.0 : res;
  if (OVERFLOW) {
    // x > 709 (0x4086280000000000) up to and including +inf gives +inf; NaNs (canonical after
    // the first fma, so n == 0) fall outside both ranges and propagate through p
    const bool overflow = (unsigned)(hi - 0x40862801) <= (0x7FF00000u - 0x40862801u);
    out = overflow ? __hiloint2double(0x7FF00000, 0) : out;
  }
  return out;
}
__device__ __forceinline__ double chi_exp(double x) { return chi_exp<true>(x, load_expk()); }

// Table-driven variant used by the hot kernels: exp(x) = 2^m * T[j] * exp(r) with
// n = rint(x * 64/ln2) = 64 m + j, |r| <= ln2/128, T[j] = 2^(j/64) in shared memory and
// exp(r) - 1 = r + r^2 g(r), g of degree 3 (fit error 5e-18).  10 FP64-pipe instructions and a
// dependent chain of 9 instead of 16 and 15: the pipe is the bound of these kernels and the
// chain length sets how much parallelism is needed to fill it.  Result within ~1 ulp.
__constant__ double c_expt[8] = {
    92.33248261689366,        // 0  64/ln2
    6755399441055744.0,       // 1  1.5 * 2^52
    -0.010830424667801708,    // 2  -(ln2/64) hi (29 significant bits: n * hi is exact)
    -2.8447437476627285e-11,  // 3  -(ln2/64) lo
    0.008333339386755335,     // 4  g3
    0.041666709040625166,     // 5  g2
    0.1666666666666436,       // 6  g1
    0.4999999999998384};      // 7  g0
__constant__ double c_exp_table[64] = {
    1.0, 1.0108892860517005, 1.0218971486541166, 1.0330248790212284,
    1.0442737824274138, 1.0556451783605572, 1.0671404006768237, 1.0787607977571199,
    1.0905077326652577, 1.102382583307841, 1.1143867425958924, 1.1265216186082418,
    1.1387886347566916, 1.1511892299529827, 1.1637248587775775, 1.1763969916502812,
    1.189207115002721, 1.202156731452703, 1.215247359980469, 1.22848053610687,
    1.241857812073484, 1.255380757024691, 1.2690509571917332, 1.2828700160787783,
    1.2968395546510096, 1.3109612115247644, 1.3252366431597413, 1.339667524053303,
    1.3542555469368927, 1.3690024229745905, 1.383909881963832, 1.3989796725383112,
    1.4142135623730951, 1.42961333839197, 1.4451808069770467, 1.460917794180647,
    1.4768261459394993, 1.4929077282912648, 1.5091644275934228, 1.5255981507445384,
    1.5422108254079407, 1.559004400237837, 1.5759808451078865, 1.593142151342267,
    1.6104903319492543, 1.6280274218573478, 1.645755478153965, 1.6636765803267364,
    1.681792830507429, 1.7001063537185235, 1.718619298122478, 1.7373338352737062,
    1.7562521603732995, 1.7753764925265212, 1.7947090750031072, 1.8142521755003989,
    1.8340080864093424, 1.8539791250833855, 1.8741676341103, 1.8945759815869656,
    1.9152065613971474, 1.9360617934922943, 1.9571441241754002, 1.978456026387951};

struct ExpT {
  double k[8];
  const double* T;  // shared-memory copy of c_exp_table
};
__device__ __forceinline__ ExpT load_expt(const double* shared_table) {
  ExpT e;
#pragma unroll
  for (int i = 0; i < 8; ++i) e.k[i] = c_expt[i];
  e.T = shared_table;
  return e;
}
// every thread of the block must call this before any early exit (block barrier inside)
__device__ __forceinline__ void fill_exp_table(double* shared_table) {
  if (threadIdx.x < 64) shared_table[threadIdx.x] = c_exp_table[threadIdx.x];
  __syncthreads();
}

template <bool OVERFLOW = true>
__device__ __forceinline__ double chi_exp(double x, const ExpT& c) {
  const int hi = __double2hiint(x);
  const double t = fma(x, c.k[0], c.k[1]);
  const double nf = t - c.k[1];
  double r = fma(nf, c.k[2], x);
  r = fma(nf, c.k[3], r);
  const int n = __double2loint(t);
  const double Tj = c.T[n & 63];
  double q = fma(c.k[4], r, c.k[5]);
  q = fma(q, r, c.k[6]);
  q = fma(q, r, c.k[7]);
  const double e = fma(r * r, q, r);  // exp(r) - 1
  const double p = fma(Tj, e, Tj);
  const double res = __hiloint2double(__double2hiint(p) + ((n >> 6) << 20), __double2loint(p));
  const bool underflow = (unsigned)(hi - 0xC0862001) <= (0xFFF00000u - 0xC0862001u);
  double out = underflow ? 0.0 : res;
  if (OVERFLOW) {
    const bool overflow = (unsigned)(hi - 0x40862801) <= (0x7FF00000u - 0x40862801u);
    out = overflow ? __hiloint2double(0x7FF00000, 0) : out;
  }
  return out;
}

#ifndef CHI_EXP_POLY
typedef ExpT ExpCtx;   // hot kernels: table-driven exp
#define CHI_LOAD_EXPCTX(tab) load_expt(tab)
#else
typedef ExpK ExpCtx;   // -DCHI_EXP_POLY: pure-polynomial exp (A/B builds)
#define CHI_LOAD_EXPCTX(tab) load_expk()
#endif

// 1/x for normal x > 0 (x = d^2 + W^2/4 of the Lorentzian): MUFU.RCP64H seed + two
// Newton steps, ~1 ulp, no slow path.
__device__ __forceinline__ double chi_rcp(double x) {
  double y;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
  double e = fma(-x, y, 1.0);
  y = fma(y, e, y);
  e = fma(-x, y, 1.0);
  y = fma(y, e, y);
  return y;
}

// Per-cloud constants live in shared memory (broadcast reads); the volatile asm keeps the
// compiler from hoisting all of them into registers, which would cost occupancy.
__device__ __forceinline__ void lds2(const double* p, double& a, double& b) {
  unsigned addr = (unsigned)__cvta_generic_to_shared(p);
  asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(a), "=d"(b) : "r"(addr));
}

__device__ __forceinline__ double warp_sum(double x) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) x += __shfl_xor_sync(CHI_FULL, x, o);
  return x;
}

// ---------------------------------------------------------------------------
// argument block of the channel kernels
// ---------------------------------------------------------------------------
struct ChanArgs {
  int64_t B;
  int S;            // channels of this dataset
  int n_base;       // baseline terms
  int n_chunks;     // channel chunks per draw
  int chunk_len;    // channels per chunk (multiple of 32)
  int mom_stride;   // doubles per (draw, chunk) record
  int axis_dir;     // +1 / -1: spectral axis strictly ascending / descending, 0: neither
  double bg;
  const double* __restrict__ vax;     // [S]
  const double* __restrict__ yeff;    // [n_sl][S]   brightness - baseline offset
  const double* __restrict__ wgt;     // [n_sl][S]   1/noise^2
  const double* __restrict__ basis;   // [n_base][S]
  const double* __restrict__ coeff;   // [B][n_base] or null
  const int32_t* __restrict__ sl;     // [B] or null
  const double* __restrict__ gbar;    // [B][S] cotangent (vjp mode) or null (logp mode)
  const double* __restrict__ cc;      // [B][N][12]
  double* __restrict__ mom;           // [B][n_chunks][mom_stride]
};

// record layout: [0]=sum -0.5 w r^2, [1..n_base]=d/dcoeff, then per cloud NE/NA moments
//   emission:   M0 M1 M2 (sum g E d^j), aTs, af, [Q1 Q2 Q2d]
//   absorption: M0 M1 M2, [Q1 Q2 Q2d]
template <bool PV> struct EmLayout { static constexpr int NE = PV ? 8 : 5; };
template <bool PV> struct AbLayout { static constexpr int NA = PV ? 6 : 3; };

// resident blocks per SM asked of the compiler: accumulators + per-cloud state are the
// register floor of a variant, everything else is traded against occupancy
#ifndef CHI_BLOCKS_BIAS
#define CHI_BLOCKS_BIAS 0  // build-time experiment knob: +/- resident blocks per SM
#endif
constexpr int clamp_blocks(int b) { return b < 1 ? 1 : (b > 6 ? 6 : b); }
constexpr int em_min_blocks(int N, bool pv) {
  const int regs = 2 * ((pv ? 8 : 5) * N + (pv ? 4 : 3) * N) + 64;
  return clamp_blocks((regs <= 128 ? 4 : (regs <= 168 ? 3 : 2)) + CHI_BLOCKS_BIAS) * 4 / CHI_WARPS_PER_BLOCK;
}
constexpr int ab_min_blocks(int N, bool pv) {
  const int regs = 2 * ((pv ? 6 : 3) * N + (pv ? 3 : 2) * N) + 56;
  return clamp_blocks((regs <= 128 ? 4 : (regs <= 168 ? 3 : 2)) + CHI_BLOCKS_BIAS) * 4 / CHI_WARPS_PER_BLOCK;
}

// first two baseline terms are kept in registers (degree 0/1 is the common case);
// further terms go through shared memory
#define CHI_BASE_REG 2

// ---------------------------------------------------------------------------
// k_moments_em
//   physics.py:294-309 (profile), emission_model.py:148 (tau), physics.py:345-365
//   (radiative transfer), pm.Normal logp, and the adjoint of all of it.
//   shared per cloud, fetched in pairs (the Gaussian case touches the first three):
//   [c, -k] [-AG, f] [f*Ts, Ts] [-AL, h]
// ---------------------------------------------------------------------------
template <int N, bool PV>
__global__ void __launch_bounds__(CHI_BLOCK, em_min_blocks(N, PV)) k_moments_em(ChanArgs a) {
  constexpr int NE = EmLayout<PV>::NE;
  __shared__ __align__(16) double s_c[CHI_WARPS_PER_BLOCK][N][8];
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

  if (lane < N) {
    const double* p = a.cc + (b * N + lane) * CHI_CC_STRIDE;
    const double f = p[CC_F], ts = p[CC_TS];
    double* d = s_c[warp][lane];
    d[0] = p[CC_C]; d[1] = -p[CC_KE]; d[2] = -p[CC_AGE]; d[3] = f; d[4] = f * ts; d[5] = ts;
    d[6] = -p[CC_ALE]; d[7] = p[CC_HE];
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
  const double* __restrict__ V = a.vax;

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
  for (int s = chunk * a.chunk_len + lane; s < s_end; s += 32) {
    const double v = V[s];
    double p0 = 0.0, p1 = 0.0, base = 0.0;
    if (nb > 0) { p0 = bas0[s]; base = beta0 * p0; }
    if (nb > 1) { p1 = bas1[s]; base = fma(beta1, p1, base); }
    for (int i = CHI_BASE_REG; i < nb; ++i) base = fma(s_beta[warp][i], a.basis[(int64_t)i * S + s], base);

    // ---- forward, nearest cloud first ----
    double om[N], Pn[N], E[N], q[N];
    double P = 1.0, T = 0.0;
#pragma unroll
    for (int n = 0; n < N; ++n) {
      double c0, nk, nAG, f, fTs, Ts;
      lds2(&sc[n][0], c0, nk);
      lds2(&sc[n][2], nAG, f);
      lds2(&sc[n][4], fTs, Ts);
      const double d = v - c0, d2 = d * d;
      E[n] = chi_exp<false>(nk * d2, ek);
      double x = nAG * E[n];                    // -tau
      if (PV) {
        double nAL, h;
        lds2(&sc[n][6], nAL, h);
        q[n] = chi_rcp(d2 + h);
        x = fma(nAL, q[n], x);
      }
      om[n] = 1.0 - chi_exp(x, ek);               // 1 - e^-tau
      Pn[n] = P;
      T = fma(fTs * om[n], P, T);               // f Ts (1-e) x attenuation in front
      P *= fma(-f, om[n], 1.0);                 // 1 - f + f e
    }
    const double mu = (bg * P + T) - bg + base;  // ON - OFF, physics.py:362-365
    double mubar;
    if (G) {
      mubar = G[s];
    } else {
      const double r = Y[s] - mu;
      mubar = Wt[s] * r;                        // d logp / d mu
      lp = fma(-0.5 * mubar, r, lp);
    }
    if (nb > 0) gb0 = fma(mubar, p0, gb0);
    if (nb > 1) gb1 = fma(mubar, p1, gb1);
    for (int i = CHI_BASE_REG; i < nb; ++i)
      s_gb[i - CHI_BASE_REG][threadIdx.x] =
          fma(mubar, a.basis[(int64_t)i * S + s], s_gb[i - CHI_BASE_REG][threadIdx.x]);

    // ---- reverse, farthest cloud first; Q = brightness arriving behind cloud n ----
    double Q = bg;
#pragma unroll
    for (int n = N - 1; n >= 0; --n) {
      double c0, nk, nAG, f, fTs, Ts;
      lds2(&sc[n][0], c0, nk);
      lds2(&sc[n][2], nAG, f);
      lds2(&sc[n][4], fTs, Ts);
      const double Sn = Ts - Q;                 // Ts_n - Q_n
      const double x = om[n] * Pn[n];
      const double mx = mubar * x;
      acc[n][3] += mx;                          // -> d/dTs (times f)
      acc[n][4] = fma(mx, Sn, acc[n][4]);       // -> d/df
      const double g = (mubar * Sn) * (Pn[n] - x);  // taubar / f = mubar e P (Ts - Q)
      const double d = v - c0;
      const double t = g * E[n];
      const double td = t * d;
      acc[n][0] += t;
      acc[n][1] += td;
      acc[n][2] = fma(td, d, acc[n][2]);
      if (PV) {
        const double u = g * q[n];
        const double u2 = u * q[n];
        acc[n][5] += u;
        acc[n][6] += u2;
        acc[n][7] = fma(u2, d, acc[n][7]);
      }
      Q = fma(fma(-f, om[n], 1.0), Q, fTs * om[n]);
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
// k_moments_abs: absorption spectrum 1 - exp(-sum_n tau)
//   absorption_model.py:88-91, emission_absorption_model.py:84-99 + physical equivalents
//   shared per cloud: c, -k, -AG, -AL, h
// ---------------------------------------------------------------------------
template <int N, bool PV>
__global__ void __launch_bounds__(CHI_BLOCK, ab_min_blocks(N, PV)) k_moments_abs(ChanArgs a) {
  constexpr int NA = AbLayout<PV>::NA;
  __shared__ __align__(16) double s_c[CHI_WARPS_PER_BLOCK][N][6];
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

  if (lane < N) {
    const double* p = a.cc + (b * N + lane) * CHI_CC_STRIDE;
    double* d = s_c[warp][lane];
    d[0] = p[CC_C]; d[1] = -p[CC_KA]; d[2] = -p[CC_AGA]; d[3] = -p[CC_ALA]; d[4] = p[CC_HA];
    d[5] = 0.0;
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
  const double* __restrict__ V = a.vax;

  double acc[N][NA];
#pragma unroll
  for (int n = 0; n < N; ++n)
#pragma unroll
    for (int j = 0; j < NA; ++j) acc[n][j] = 0.0;
  double lp = 0.0, gb0 = 0.0, gb1 = 0.0;
  const double(*sc)[6] = s_c[warp];
  const ExpCtx ek = CHI_LOAD_EXPCTX(s_T);

  const int s_end = min(S, (chunk + 1) * a.chunk_len);
  for (int s = chunk * a.chunk_len + lane; s < s_end; s += 32) {
    const double v = V[s];
    double p0 = 0.0, p1 = 0.0, base = 0.0;
    if (nb > 0) { p0 = bas0[s]; base = beta0 * p0; }
    if (nb > 1) { p1 = bas1[s]; base = fma(beta1, p1, base); }
    for (int i = CHI_BASE_REG; i < nb; ++i) base = fma(s_beta[warp][i], a.basis[(int64_t)i * S + s], base);

    double E[N], d[N], q[N];
    double xs = 0.0;  // -sum_n tau_n
#pragma unroll
    for (int n = 0; n < N; ++n) {
      double c0, nk, nAG, nAL;
      lds2(&sc[n][0], c0, nk);
      lds2(&sc[n][2], nAG, nAL);
      d[n] = v - c0;
      const double d2 = d[n] * d[n];
      E[n] = chi_exp<false>(nk * d2, ek);
      xs = fma(nAG, E[n], xs);
      if (PV) {
        q[n] = chi_rcp(d2 + sc[n][4]);
        xs = fma(nAL, q[n], xs);
      }
    }
    const double ea = chi_exp(xs, ek);
    const double mu = (1.0 - ea) + base;
    double mubar;
    if (G) {
      mubar = G[s];
    } else {
      const double r = Y[s] - mu;
      mubar = Wt[s] * r;
      lp = fma(-0.5 * mubar, r, lp);
    }
    if (nb > 0) gb0 = fma(mubar, p0, gb0);
    if (nb > 1) gb1 = fma(mubar, p1, gb1);
    for (int i = CHI_BASE_REG; i < nb; ++i)
      s_gb[i - CHI_BASE_REG][threadIdx.x] =
          fma(mubar, a.basis[(int64_t)i * S + s], s_gb[i - CHI_BASE_REG][threadIdx.x]);

    const double g = mubar * ea;  // taubar, identical for every cloud
#pragma unroll
    for (int n = 0; n < N; ++n) {
      const double t = g * E[n];
      const double td = t * d[n];
      acc[n][0] += t;
      acc[n][1] += td;
      acc[n][2] = fma(td, d[n], acc[n][2]);
      if (PV) {
        const double u = g * q[n];
        const double u2 = u * q[n];
        acc[n][3] += u;
        acc[n][4] += u2;
        acc[n][5] = fma(u2, d[n], acc[n][5]);
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
    for (int j = 0; j < NA; ++j) {
      double r = warp_sum(acc[n][j]);
      if (lane == ((n * NA + j) & 31)) om_[n * NA + j] = r;
    }
}

}  // namespace chi
