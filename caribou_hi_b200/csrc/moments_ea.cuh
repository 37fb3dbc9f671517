// Fused emission + absorption kernel for datasets that share one spectral axis.
//
// When both spectra are sampled on the same velocity grid, the convolved line width, k = 4ln2/W^2
// and h = W^2/4 of a cloud are the same for the two datasets and so is exp(-k d^2): one pass over
// the channels evaluates it once and feeds both the radiative transfer (emission) and the summed
// optical depth (absorption), cutting the absorption side from ~25 to ~9 FP64 instructions per
// (draw, cloud, channel).  The gradients with respect to the line centre and k
// (and h) are accumulated through COMBINED moments
//     M1c = sum (f AG_e g_e + AG_a g_a) E d,   M2c = sum (...) E d^2,   [Q2c, Q2dc likewise with AL, q^2]
// so the fused record needs 6 (Gaussian) / 10 (pseudo-Voigt) sums per cloud instead of 5+3 / 8+6:
//     M0e = sum g_e E, M0a = sum g_a E, M1c, M2c, aTs, af, [Q1e, Q1a, Q2c, Q2dc]
// The per-channel data of the two datasets is read from one packed array (FusedArgs::chan): per
// group of 32 channels eight planes of 32 doubles -- v, p_e0, p_a0, -, y_e, w_e, y_a, w_a; p_x0 =
// first baseline basis column or 0 -- so that a warp's iteration has one address, every plane
// is an immediate offset from it and each load stays a coalesced 256-byte request (an
// array-of-structs layout costs 16 L1 wavefronts per load and saturates the LSU data pipe).  SB ("simple
// baseline") instantiations serve n_baseline <= 1 on both datasets -- the reference's default
// baseline_degree = 0 -- and carry no shared-memory baseline loops at all.
// Record per (draw, chunk): [0] = sum -0.5 w r^2 of both datasets, [1..nb_e] and [1+nb_e..] the
// baseline-coefficient gradients, then N x NF moments.  k_epilogue reads it with `fused = 1`.
#pragma once
#include "moments.cuh"

namespace chi {

struct FusedArgs {
  int64_t B;
  int S;
  int nb_e, nb_a;
  int n_real;                          // clouds of the model (<= template N, see ChanArgs::n_real)
  int staged;                          // 1: more than one sightline is loaded (survey mode)
  int n_chunks, chunk_len, mom_stride;
  double bg;
  const double* __restrict__ chan;     // [n_sl][ceil(S/32)][8][32] packed channel data (see above)
  const double* __restrict__ basis_e;
  const double* __restrict__ coeff_e;
  const double* __restrict__ gbar_e;
  const double* __restrict__ basis_a;
  const double* __restrict__ coeff_a;
  const double* __restrict__ gbar_a;
  int vjp;                             // 1: cotangents gbar_* instead of residuals (null = zeros)
  const int32_t* __restrict__ sl;
  const double* __restrict__ cc;
  double* __restrict__ mom;
  unsigned long long* n_marked;        // EA_FIXUP: += the draws it evaluated (null: not counted)
};

template <bool PV> struct FusedLayout { static constexpr int NF = PV ? 10 : 6; };

constexpr int ea_min_blocks(int N, bool pv) {
  const int regs = 2 * ((pv ? 10 : 6) * N + (pv ? 4 : 3) * N) + 56;
  return clamp_blocks((regs <= 128 ? 4 : (regs <= 168 ? 3 : 2)) + CHI_BLOCKS_BIAS) * 4 / CHI_WARPS_PER_BLOCK;
}

// shared per cloud, fetched in pairs; ' = times log2(e) (arguments of chi_exp2).  The Gaussian case
// reads A B C forward and A C D in reverse:
//   A [c, -k']  B [-AG_e', f*Ts]  C [f, -AG_a']  D [Ts, -f AG_e']  E [-AL_e', -AL_a']  F [h, -f AL_e']
__device__ __forceinline__ void ldg2(const double* p, double& x, double& y) {
  const double2 t = __ldg(reinterpret_cast<const double2*>(p));
  x = t.x;
  y = t.y;
}

// STAGED: see the bulk-async helpers in moments.cuh.
// VJP = false: logp mode only (a.vjp == 0 is the caller's promise): the cotangent pointers and the
// mode test leave the channel loop, which is short of registers.
// MODE: 2^y of a NEGATIVE optical depth can overflow, and the +inf branch of chi_exp2 costs two selects, an
// add and a compare per exponential (20 integer-pipe instructions per channel at N = 4, 2.2 % of the
// kernel).  The cloud map marks the clouds whose depths are negative (cc[..][CC_CAREFUL]); a large batch
// then runs as two launches side by side: EA_TAME over every draw without that branch (a marked draw
// leaves at once), and EA_FIXUP, one wave of resident blocks that scan the marks and evaluate the marked
// draws with it (normally none: the scan takes microseconds next to the other launch).  The scan is
// built for FEW marks: a marked draw is one warp's work for ~34 us whatever the rest of the machine does, and the
// static partition (32 draws per warp and pass) balances a batch FULL of marked draws badly -- 64 draws on some
// warps, 32 on others: 2.45 against 1.47 ms.  So EA_FIXUP counts the draws it evaluated and the host side goes
// back to EA_ALL while recent calls had more than 0.5 % of them (two_pass_ok in chi.cu).  EA_ALL is the single
// launch with the branch, also for small batches.
// (enum { EA_ALL, EA_TAME, EA_FIXUP }: moments.cuh)

template <int N, bool PV, bool SB, bool STAGED = false, bool VJP = true, int MODE = EA_ALL>
__global__ void __launch_bounds__(CHI_BLOCK, ea_min_blocks(N, PV)) k_moments_ea(FusedArgs a) {
  constexpr int NF = FusedLayout<PV>::NF;
  constexpr bool OVF = MODE != EA_TAME;
  static_assert(!(STAGED && MODE == EA_FIXUP), "the scan form reads the channel data through L1");
  __shared__ __align__(16) double s_c[CHI_WARPS_PER_BLOCK][N][12];
  __shared__ double s_gb[SB ? 1 : 2 * CHI_MAX_BASELINE][SB ? 1 : CHI_BLOCK];
  __shared__ double s_beta[CHI_WARPS_PER_BLOCK][2 * CHI_MAX_BASELINE];
  __shared__ double s_T[CHI_EXP_TABLE];
  __shared__ __align__(128) double s_stage[STAGED ? CHI_WARPS_PER_BLOCK : 1][2][STAGED ? 256 : 2];
  __shared__ __align__(8) unsigned long long s_bar[STAGED ? CHI_WARPS_PER_BLOCK : 1][2];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (STAGED && lane == 0) {
    mbar_init((unsigned)__cvta_generic_to_shared(&s_bar[warp][0]));
    mbar_init((unsigned)__cvta_generic_to_shared(&s_bar[warp][1]));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  fill_exp_table(s_T);  // contains the block barrier

  int it = 0;  // STAGED: groups fetched so far (buffer and mbarrier phase)
  auto one_item = [&](const int64_t item) {
    const int64_t b = item / a.n_chunks;
    const int chunk = (int)(item - b * a.n_chunks);
    const int nbe = a.nb_e, nba = a.nb_a, S = a.S;

    bool marked = false;
    if (lane >= a.n_real && lane < N) {  // null cloud
      double* d = s_c[warp][lane];
      for (int i = 0; i < 12; ++i) d[i] = 0.0;
      d[1] = -1.0; d[10] = 1.0;
    }
    if (lane < a.n_real) {
      const double* p = a.cc + (b * a.n_real + lane) * CHI_CC_STRIDE;
      const double f = p[CC_F], ts = p[CC_TS];
      double* d = s_c[warp][lane];
      const double age = -CHI_LOG2E * p[CC_AGE], ale = -CHI_LOG2E * p[CC_ALE];
      d[0] = p[CC_C]; d[1] = -CHI_LOG2E * p[CC_KE]; d[2] = age; d[3] = f * ts; d[4] = f;
      d[5] = -CHI_LOG2E * p[CC_AGA]; d[6] = ts; d[7] = f * age; d[8] = ale; d[9] = -CHI_LOG2E * p[CC_ALA];
      d[10] = p[CC_HE]; d[11] = f * ale;
      marked = p[CC_CAREFUL] != 0.0;
    }
    if (MODE == EA_TAME && __any_sync(0xffffffffu, marked)) return;  // EA_FIXUP's draw
    if (lane < nbe) s_beta[warp][lane] = a.coeff_e ? a.coeff_e[b * nbe + lane] : 0.0;
    if (lane < nba) s_beta[warp][CHI_MAX_BASELINE + lane] = a.coeff_a ? a.coeff_a[b * nba + lane] : 0.0;
    if (!SB) {
      for (int i = 1; i < nbe; ++i) s_gb[i][threadIdx.x] = 0.0;
      for (int i = 1; i < nba; ++i) s_gb[CHI_MAX_BASELINE + i][threadIdx.x] = 0.0;
    }
    __syncwarp();
    // the first baseline term of each dataset (degree 0) stays in registers
    const double be0 = nbe > 0 ? s_beta[warp][0] : 0.0, ba0 = nba > 0 ? s_beta[warp][CHI_MAX_BASELINE] : 0.0;
    double gbe0 = 0.0, gba0 = 0.0;

    const int sl = a.sl ? a.sl[b] : 0;
    const double* __restrict__ ch = a.chan + (int64_t)sl * ((S + 31) >> 5) * 256 + lane;
    const double* __restrict__ Ge = (VJP && a.gbar_e) ? a.gbar_e + b * S : nullptr;
    const double* __restrict__ Ga = (VJP && a.gbar_a) ? a.gbar_a + b * S : nullptr;

    double acc[N][NF];
  #pragma unroll
    for (int n = 0; n < N; ++n)
  #pragma unroll
      for (int j = 0; j < NF; ++j) acc[n][j] = 0.0;
    double lp = 0.0;
    const double bg = a.bg;
    const double(*sc)[12] = s_c[warp];
    const ExpCtx ek = CHI_LOAD_EXPCTX(s_T);

    const int s_end = min(S, (chunk + 1) * a.chunk_len);
    const unsigned stage_addr = (unsigned)__cvta_generic_to_shared(&s_stage[STAGED ? warp : 0][0][0]);
    const unsigned bar_addr = (unsigned)__cvta_generic_to_shared(&s_bar[STAGED ? warp : 0][0]);
    if (STAGED && lane == 0 && chunk * a.chunk_len < s_end)
      tma_fetch<2048>(ch - lane + (int64_t)((chunk * a.chunk_len) >> 5) * 256, stage_addr, bar_addr);
    // STAGED: the whole warp runs every group (lanes past a ragged end carry zero weight) because
    // all lanes take part in the buffer hand-over
    for (int s = chunk * a.chunk_len + lane; (STAGED ? s - lane : s) < s_end; s += 32) {
      const double* __restrict__ cs = ch + (int64_t)(s >> 5) * 256;  // lane offset is in ch
      const int sx = STAGED ? min(s, s_end - 1) : s;
      const double* __restrict__ sg = cs;
      double v, pe0, pa0;
      if (STAGED) {
        const int buf = it & 1;
        __syncwarp();  // every lane is done reading the other buffer
        if (lane == 0 && s - lane + 32 < s_end)
          tma_fetch<2048>(cs - lane + 256, stage_addr + (buf ^ 1) * 2048, bar_addr + (buf ^ 1) * 8);
        mbar_wait(bar_addr + buf * 8, (it >> 1) & 1);
        sg = &s_stage[warp][buf][lane];
        ++it;
        v = sg[0]; pe0 = sg[32]; pa0 = sg[64];
      } else {
        v = __ldg(cs); pe0 = __ldg(cs + 32); pa0 = __ldg(cs + 64);
      }
      double base_e = be0 * pe0, base_a = ba0 * pa0;  // p_x0 is 0 without a baseline
      if (!SB) {
        for (int i = 1; i < nbe; ++i) base_e = fma(s_beta[warp][i], a.basis_e[(int64_t)i * S + s], base_e);
        for (int i = 1; i < nba; ++i)
          base_a = fma(s_beta[warp][CHI_MAX_BASELINE + i], a.basis_a[(int64_t)i * S + s], base_a);
      }

      // ---- forward: one exp(-k d^2) per cloud feeds both datasets ----
      double om[N], Pn[N], E[N], q[N];
      double P = 1.0, T = 0.0, xs = 0.0;  // xs = -sum_n tau_a
  #pragma unroll
      for (int n = 0; n < N; ++n) {
        double c0, nk, nAGe, nAGa, f, fTs;
        lds2(&sc[n][0], c0, nk);
        lds2(&sc[n][2], nAGe, fTs);
        lds2(&sc[n][4], f, nAGa);
        const double d = v - c0, d2 = d * d;
        E[n] = chi_exp2_prod<false>(nk, d2, ek);
        xs = fma(nAGa, E[n], xs);
        if (PV) {
          double x = nAGe * E[n];
          double h, fALe, nALe, nALa;
          lds2(&sc[n][8], nALe, nALa);
          lds2(&sc[n][10], h, fALe);
          q[n] = chi_rcp(d2 + h);
          x = fma(nALe, q[n], x);
          xs = fma(nALa, q[n], xs);
          om[n] = 1.0 - chi_exp2<OVF>(x, ek);
        } else {
          om[n] = 1.0 - chi_exp2_prod<OVF, OVF>(nAGe, E[n], ek);
        }
        Pn[n] = P;
        const double u = om[n] * P;
        T = fma(fTs, u, T);
        P = fma(-f, u, P);
      }
      const double mu_e = (bg * P + T) - bg + base_e;   // physics.py:362-365
      const double ea = chi_exp2<OVF, OVF>(xs, ek);
      const double mu_a = (1.0 - ea) + base_a;          // absorption_model.py:91
      double mbe, mba;
      if (VJP && a.vjp) {
        mbe = Ge ? Ge[sx] : 0.0;
        mba = Ga ? Ga[sx] : 0.0;
      } else {
        double ye, we, ya, wa;
        if (STAGED) {
          ye = sg[128]; we = sg[160]; ya = sg[192]; wa = sg[224];
        } else {
          ye = __ldg(cs + 128); we = __ldg(cs + 160); ya = __ldg(cs + 192); wa = __ldg(cs + 224);
        }
        const double re = ye - mu_e, ra = ya - mu_a;
        mbe = we * re;
        mba = wa * ra;
        if (STAGED && s >= s_end) { mbe = 0.0; mba = 0.0; }  // padded lanes of a ragged end
        lp = fma(mbe, re, lp);                          // -0.5 applied once at the end
        lp = fma(mba, ra, lp);
      }
      if (STAGED && s >= s_end) { mbe = 0.0; mba = 0.0; }
      gbe0 = fma(mbe, pe0, gbe0);
      gba0 = fma(mba, pa0, gba0);
      if (!SB) {
        for (int i = 1; i < nbe; ++i)
          s_gb[i][threadIdx.x] = fma(mbe, a.basis_e[(int64_t)i * S + s], s_gb[i][threadIdx.x]);
        for (int i = 1; i < nba; ++i)
          s_gb[CHI_MAX_BASELINE + i][threadIdx.x] =
              fma(mba, a.basis_a[(int64_t)i * S + s], s_gb[CHI_MAX_BASELINE + i][threadIdx.x]);
      }
      const double ga = mba * ea;  // absorption taubar (same for every cloud)

      // ---- reverse through the transfer; combined moments for the shared profile ----
      double Q = bg;
  #pragma unroll
      for (int n = N - 1; n >= 0; --n) {
        double c0, nk, f, nAGa, Ts, fAGe;
        lds2(&sc[n][0], c0, nk);
        lds2(&sc[n][4], f, nAGa);
        lds2(&sc[n][6], Ts, fAGe);
        const double Sn = Ts - Q;
        const double x = om[n] * Pn[n];
        const double mx = mbe * x;
        acc[n][4] += mx;                                // aTs
        acc[n][5] = fma(mx, Sn, acc[n][5]);             // af
        const double ge = (mbe * Sn) * (Pn[n] - x);     // emission taubar / f
        const double d = v - c0;
        acc[n][0] = fma(ge, E[n], acc[n][0]);           // M0e
        acc[n][1] = fma(ga, E[n], acc[n][1]);           // M0a
        // (f AG_e ge + AG_a ga) log2(e): the constants are stored negated and scaled
        const double gc = -fma(fAGe, ge, nAGa * ga);
        const double t = gc * E[n];
        const double td = t * d;
        acc[n][2] += td;                                // M1c
        acc[n][3] = fma(td, d, acc[n][3]);              // M2c
        if (PV) {
          acc[n][6] = fma(ge, q[n], acc[n][6]);         // Q1e
          acc[n][7] = fma(ga, q[n], acc[n][7]);         // Q1a
          double nALe, nALa, h, fALe;
          lds2(&sc[n][8], nALe, nALa);
          lds2(&sc[n][10], h, fALe);
          const double gl = -fma(fALe, ge, nALa * ga);
          const double u2 = gl * q[n] * q[n];
          acc[n][8] += u2;                              // Q2c
          acc[n][9] = fma(u2, d, acc[n][9]);            // Q2dc
        }
        Q = fma(om[n], f * Sn, Q);                      // (1 - f om) Q + f Ts om
      }
    }

    double* out = a.mom + item * a.mom_stride;
    if (!SB) {
      for (int i = 1; i < nbe; ++i) {
        double g = warp_sum(s_gb[i][threadIdx.x]);
        if (lane == 0) out[1 + i] = g;
      }
      for (int i = 1; i < nba; ++i) {
        double g = warp_sum(s_gb[CHI_MAX_BASELINE + i][threadIdx.x]);
        if (lane == 0) out[1 + nbe + i] = g;
      }
    }
    // logp, the two register-held baseline gradients and the N * NF moments: one folded reduction
    typedef WarpFold<3 + N * NF> Fold;
    double v[3 + N * NF];
    v[0] = lp; v[1] = gbe0; v[2] = gba0;
  #pragma unroll
    for (int n = 0; n < N; ++n)
  #pragma unroll
      for (int j = 0; j < NF; ++j) v[3 + n * NF + j] = acc[n][j];
    Fold::run(v, lane);
    double* om_ = out + 1 + nbe + nba;
  #pragma unroll
    for (int i = 0; i < Fold::OUT; ++i) {
      const int idx = Fold::index(i, lane);
      if (idx >= 3) {
        const int j = (idx - 3) % NF;
        // combined moments carry log2(e)
        om_[idx - 3] = (j == 2 || j == 3 || j == 8 || j == 9) ? v[i] * CHI_LN2 : v[i];
      } else if (idx == 0) {
        out[0] = -0.5 * v[i];
      } else if (idx == 1) {
        if (nbe > 0) out[1] = v[i];
      } else if (idx == 2) {
        if (nba > 0) out[1 + nbe] = v[i];
      }
    }
  };
  if (MODE != EA_FIXUP) {
    const int64_t item = (int64_t)blockIdx.x * CHI_WARPS_PER_BLOCK + warp;
    if (item < a.B * a.n_chunks) one_item(item);
  } else {
    // 32 draws per warp and pass: one ballot of their marks, then the marked ones chunk by chunk
    const int64_t n_warps = (int64_t)gridDim.x * CHI_WARPS_PER_BLOCK;
    for (int64_t b0 = ((int64_t)blockIdx.x * CHI_WARPS_PER_BLOCK + warp) * 32; b0 < a.B; b0 += n_warps * 32) {
      bool m = false;
      if (b0 + lane < a.B)
        for (int n = 0; n < a.n_real; ++n)
          m |= __ldg(a.cc + ((b0 + lane) * a.n_real + n) * CHI_CC_STRIDE + CC_CAREFUL) != 0.0;
      const unsigned marks = __ballot_sync(0xffffffffu, m);
      if (marks && lane == 0 && a.n_marked) atomicAdd(a.n_marked, (unsigned long long)__popc(marks));
      for (unsigned todo = marks; todo; todo &= todo - 1) {
        const int64_t b = b0 + (__ffs(todo) - 1);
        for (int chunk = 0; chunk < a.n_chunks; ++chunk) {
          one_item(b * a.n_chunks + chunk);
          __syncwarp();  // the warp's shared constants are rewritten by the next item
        }
      }
    }
  }
}

}  // namespace chi
