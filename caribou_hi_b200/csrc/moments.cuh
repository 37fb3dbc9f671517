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

// Table-driven base-2 variant used by the hot kernels.  The argument is y = x log2(e): the
// per-cloud constants (-k, -A_G, -A_L) are pre-multiplied by log2(e) when they are staged in
// shared memory, so the conversion costs nothing per channel.
//   t = y + 1.5*2^42 rounds y to a multiple of 2^-10: lo32(t) = n = 1024 m + j,
//   r = y - n/1024 (exact), |r| <= 2^-11,
//   2^y = 2^m * T[j] * (1 + r q(r)),  q the least-squares quadratic of (2^r - 1)/r on |r| <= 2^-11: its
//   residual is orthogonal to r, so the error of 2^y has ZERO MEAN over the arguments (+0.002 ulp
//   measured, rms 0.57, max 2.6 ulp).  The minimax fit used before (max 1.7 ulp) was biased by
//   +0.29 ulp, which a sum over channels with same-sign residuals turns into a coherent error: the
//   two paths from a filling factor to an optically thin line cancel to 1e-5 of either, and the
//   bias of 1 - 2^(-tau) then showed as 1e-6 of that gradient component (tools/grad_probe.py),
//   T[j] = 2^(j/1024) in shared memory (8 KB per block, filled from g_exp2_table).
// 7 FP64-pipe instructions and a dependent chain of 6; result within ~1 ulp.
//   * y < -1022 returns a value below 2^-1042 (hi word zeroed; the low word is left alone, one
//     select instead of two), i.e. zero for every use on this path;
//   * OVERFLOW: y >= 1024 and +inf give +inf; NaNs propagate through r.
__device__ const double g_exp2_table[1024] = {
#include "exp2_table.inc"
};
#define CHI_LOG2E 1.4426950408889634074
// CHI_LN2: cloud_map.cuh
#ifndef CHI_EXP_REPL
#define CHI_EXP_REPL 0
#endif
#if CHI_EXP_REPL
// A/B variant (tools/exp_variants.sh), REJECTED: 128 table entries (2^(j/128)), each replicated 16 times
// so that lane l reads bank pair l % 16 whatever its index -- no bank conflicts -- and a cubic q (one
// more FMA per exponential) for the 8x wider remainder |r| <= 2^-8.  Measured at configs[1]: fused
// kernel 1.480 -> 1.566 ms (Gaussian), 2.155 -> 2.233 ms (pseudo-Voigt): the extra FP64 instruction
// costs more than the 25 % of shared-memory wavefronts that were conflicts; the kernel is bound by
// FP64 issue, not by the LSU (profiles/r2_exp_variants.txt).
__constant__ double c_expt[5] = {
    52776558133248.0,       // 0  1.5 * 2^45: rounds to multiples of 2^-7
    0x1.62e42fefa3900p-1,   // 1..4 least-squares cubic of (2^r - 1)/r on |r| <= 2^-8
    0x1.ebfbdff82c45bp-3, 0x1.c6b096cd133cfp-5, 0x1.3b2abc975b807p-7};
#define CHI_EXP_BITS 7
#define CHI_EXP_TABLE (128 * 16)
#define CHI_EXP_NK 5
#else
__constant__ double c_expt[4] = {
    6597069766656.0,        // 0  1.5 * 2^42
    0x1.62e42fefa39efp-1,   // 1  q0 = ln2
    0x1.ebfbe02772c13p-3,   // 2  q1
    0x1.c6b08d95bd306p-5};  // 3  q2
#define CHI_EXP_BITS 10
#define CHI_EXP_TABLE 1024
#define CHI_EXP_NK 4
#endif

struct ExpT {
  double k[CHI_EXP_NK];
  unsigned T;  // shared-memory address of the table (of this lane's replica in the replicated form)
};
__device__ __forceinline__ ExpT load_expt(const double* shared_table) {
  ExpT e;
#pragma unroll
  for (int i = 0; i < CHI_EXP_NK; ++i) e.k[i] = c_expt[i];
  e.T = (unsigned)__cvta_generic_to_shared(shared_table);
#if CHI_EXP_REPL
  e.T += (threadIdx.x & 15) * 8;
#endif
  return e;
}
// every thread of the block must call this before any early exit (block barrier inside)
__device__ __forceinline__ void fill_exp_table(double* shared_table) {
#if CHI_EXP_REPL
  for (int i = threadIdx.x; i < CHI_EXP_TABLE; i += blockDim.x) shared_table[i] = g_exp2_table[(i >> 4) << 3];
#else
  for (int i = threadIdx.x; i < CHI_EXP_TABLE; i += blockDim.x) shared_table[i] = g_exp2_table[i];
#endif
  __syncthreads();
}
// table look-up and polynomial shared by the two entry points: n = lo32(t), r the remainder
__device__ __forceinline__ double chi_exp2_core(int n, double r, const ExpT& c, int& m) {
  unsigned addr;
#if CHI_EXP_REPL
  asm("mad.lo.u32 %0, %1, 128, %2;" : "=r"(addr) : "r"(n & 127), "r"(c.T));
#else
  asm("mad.lo.u32 %0, %1, 8, %2;" : "=r"(addr) : "r"(n & (CHI_EXP_TABLE - 1)), "r"(c.T));
#endif
  double Tj;
  asm("ld.shared.f64 %0, [%1];" : "=d"(Tj) : "r"(addr));
#if CHI_EXP_REPL
  double q = fma(c.k[4], r, c.k[3]);
  q = fma(q, r, c.k[2]);
#else
  double q = fma(c.k[3], r, c.k[2]);
#endif
  q = fma(q, r, c.k[1]);
  const double Tr = Tj * r;
  m = n >> CHI_EXP_BITS;
  return fma(Tr, q, Tj);
}
// y < -1022 (result below 2^-1042): negative doubles order like unsigned ints on the high word; the
// range stops at -inf (0xFFF00000) so that NaNs with the sign bit set still propagate.
// CHI_EXP_U (A/B variant, REJECTED): one unsigned compare, hi >= 0xC08FF001.  0.6 % faster (fused kernel
// 1.480 -> 1.471 ms at configs[1], profiles/r2_exp_variants.txt) but WRONG: a NaN optical depth reaches
// here with its sign bit set (the constants are stored negated: -(log2 e * NaN)), the compare takes it for
// an underflow and exp(-tau) becomes 0 -- a NaN draw of the reference came out as a finite absorption
// spectrum (caught by tests/graph_worker.py on the reference's own NaN draws; regression test
// tests/test_gpu_parity.py::test_nan_inputs_propagate).
#ifndef CHI_EXP_U
#define CHI_EXP_U 0
#endif
__device__ __forceinline__ bool chi_exp2_underflow(int hi) {
#if CHI_EXP_U
  return (unsigned)hi >= 0xC08FF001u;
#else
  return (unsigned)(hi - 0xC08FF001) <= (0xFFF00000u - 0xC08FF001u);
#endif
}

// 2^y
// UNDERFLOW = false: the caller knows y > -1022 (or NaN), the zero select is left out
template <bool OVERFLOW = true, bool UNDERFLOW = true>
__device__ __forceinline__ double chi_exp2(double y, const ExpT& c) {
  const int hi = __double2hiint(y);
  const double t = y + c.k[0];
  const double nf = t - c.k[0];
  const double r = y - nf;
  int m;
  const double p = chi_exp2_core(__double2loint(t), r, c, m);
  int ph;
  asm("mad.lo.s32 %0, %1, 1048576, %2;" : "=r"(ph) : "r"(m), "r"(__double2hiint(p)));
  // the range stops at -inf (0xFFF00000) so that NaNs still propagate through r and p
  if (UNDERFLOW) ph = chi_exp2_underflow(hi) ? 0 : ph;
  double out = __hiloint2double(ph, __double2loint(p));
  if (OVERFLOW) {
    // y >= 1024 (0x4090000000000000) up to and including +inf gives +inf
    const bool overflow = (unsigned)(hi - 0x40900000) <= (0x7FF00000u - 0x40900000u);
    out = overflow ? __hiloint2double(0x7FF00000, 0) : out;
  }
  return out;
}

// 2^(a b): the product never materialises -- t = fma(a, b, magic) rounds a b to a multiple of 2^-10
// in one step and r = fma(a, b, -nf) is the exact remainder, one FP64 instruction fewer than
// chi_exp2(a * b); the range tests read the rounded argument nf instead of y.
template <bool OVERFLOW = true, bool UNDERFLOW = true>
__device__ __forceinline__ double chi_exp2_prod(double a, double b, const ExpT& c) {
  const double t = fma(a, b, c.k[0]);
  const double nf = t - c.k[0];
  const double r = fma(a, b, -nf);
  const int hi = __double2hiint(nf);
  int m;
  const double p = chi_exp2_core(__double2loint(t), r, c, m);
  int ph;
  asm("mad.lo.s32 %0, %1, 1048576, %2;" : "=r"(ph) : "r"(m), "r"(__double2hiint(p)));
  if (UNDERFLOW) ph = chi_exp2_underflow(hi) ? 0 : ph;
  double out = __hiloint2double(ph, __double2loint(p));
  if (OVERFLOW) {
    const bool overflow = (unsigned)(hi - 0x40900000) <= (0x7FF00000u - 0x40900000u);
    out = overflow ? __hiloint2double(0x7FF00000, 0) : out;
  }
  return out;
}

typedef ExpT ExpCtx;
#define CHI_LOAD_EXPCTX(tab) load_expt(tab)

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

// Warp totals of M per-lane values with sum_k ceil(M / 2^k) (~M) exchanges instead of 5 M: in each of
// the five butterfly steps a lane hands one half of its values to its partner and keeps, and
// completes, the other half.  Afterwards lane L holds, in v[i] for i < OUT, the total of the
// original value number index(i, L) (-1: nothing).  Fixed summation tree: deterministic.
template <int M>
struct WarpFold {
  static constexpr int N1 = (M + 1) / 2, N2 = (N1 + 1) / 2, N3 = (N2 + 1) / 2, N4 = (N3 + 1) / 2, OUT = (N4 + 1) / 2;
  template <int NIN, int MASK>
  static __device__ __forceinline__ void step(double (&v)[M], int lane) {
    constexpr int H = (NIN + 1) / 2;
    const bool up = (lane & MASK) != 0;
#pragma unroll
    for (int i = 0; i < H; ++i) {
      const double lo = v[i], hi = (i + H < NIN) ? v[i + H] : 0.0;
      const double send = up ? lo : hi, keep = up ? hi : lo;
      v[i] = keep + __shfl_xor_sync(CHI_FULL, send, MASK);
    }
  }
  static __device__ __forceinline__ void run(double (&v)[M], int lane) {
    step<M, 16>(v, lane);
    step<N1, 8>(v, lane);
    step<N2, 4>(v, lane);
    step<N3, 2>(v, lane);
    step<N4, 1>(v, lane);
  }
  static __device__ __forceinline__ int index(int i, int lane) {
    int p = i;
    bool ok = true;
    if (lane & 1) { p += OUT; ok = ok && p < N4; }
    if (lane & 2) { p += N4; ok = ok && p < N3; }
    if (lane & 4) { p += N3; ok = ok && p < N2; }
    if (lane & 8) { p += N2; ok = ok && p < N1; }
    if (lane & 16) { p += N1; ok = ok && p < M; }
    return ok ? p : -1;
  }
};

// --- bulk async copy (TMA, 1-D) + mbarrier helpers of the STAGED channel pipeline -----------
// STAGED instantiations (survey mode, one spectrum per sightline): warps of a block read different
// sightlines, so the channel data has no L1 reuse and every iteration's loads would be L2 / HBM
// latency on the critical path.  Each warp then streams its sightline through a double buffer in
// shared memory: one lane issues a bulk async copy (cp.async.bulk, completion on an mbarrier) of
// the NEXT channel group while the current one is consumed -- no registers are held across the
// load latency.  (With one shared spectrum the L1-cached loads are faster: the bulk copies bypass
// L1 and would re-read every group from L2 for every warp.)
__device__ __forceinline__ void mbar_init(unsigned bar) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar) : "memory");
}
// one lane: expect `bytes` on `bar`, then copy one packed channel group global -> shared
template <int BYTES>
__device__ __forceinline__ void tma_fetch(const double* src, unsigned dst, unsigned bar) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "n"(BYTES) : "memory");
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %3, [%2];" ::"r"(dst),
      "l"(src), "r"(bar), "n"(BYTES)
      : "memory");
}
// all lanes: wait for the phase with the given parity; a protocol error traps instead of hanging
__device__ __forceinline__ void mbar_wait(unsigned bar, unsigned parity) {
  unsigned ok = 0;
  for (int spin = 0; spin < (1 << 22); ++spin) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(bar), "r"(parity)
        : "memory");
    if (ok) return;
  }
  __trap();
}

// ---------------------------------------------------------------------------
// argument block of the channel kernels
// ---------------------------------------------------------------------------
struct ChanArgs {
  int64_t B;
  int S;            // channels of this dataset
  int n_base;       // baseline terms
  int staged;       // 1: more than one sightline is loaded (survey mode)
  int n_real;       // clouds of the model; the template N may be larger (N > 8 runs on the 12- and
                    // 16-cloud instantiations, padded with null clouds: tau = 0, f = 0)
  int n_chunks;     // channel chunks per draw
  int chunk_len;    // channels per chunk (multiple of 32)
  int mom_stride;   // doubles per (draw, chunk) record
  double bg;
  // packed channel data [n_sl][ceil(S/32)][4][32]: per group of 32 channels the planes v, p_0
  // (first baseline basis column or 0), y (brightness - baseline offset), w (1/noise^2): one
  // address per warp iteration, immediate plane offsets, coalesced 256-byte requests
  const double* __restrict__ chan;
  const double* __restrict__ basis;   // [n_base][S]
  const double* __restrict__ coeff;   // [B][n_base] or null
  const int32_t* __restrict__ sl;     // [B] or null
  const double* __restrict__ gbar;    // [B][S] cotangent (vjp mode) or null (logp mode)
  const double* __restrict__ cc;      // [B][N][12]
  double* __restrict__ mom;           // [B][n_chunks][mom_stride]
  unsigned long long* n_marked;        // EA_FIXUP: += the draws it evaluated (null: not counted)
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
//   shared per cloud, fetched in pairs (the Gaussian case touches the first three forward and
//   the first and third in reverse); k' A' = k A log2(e) feed chi_exp2:
//   [c, -k'] [-AG', f*Ts] [f, Ts] [-AL', h]
// ---------------------------------------------------------------------------
// SB ("simple baseline") instantiations serve n_base <= 1 -- the reference's default
// baseline_degree = 0 -- and carry no baseline loops.
// MODE: see moments_ea.cuh -- EA_ALL one launch with the range selects of exp(-tau) in it; EA_TAME every draw the
// cloud map did not mark, without them; EA_FIXUP a wave of resident blocks that scans the marks and evaluates the
// marked draws.  (Emission kernel, two axes: 1.197 -> 1.115 ms Gaussian, 1.517 -> 1.457 ms pseudo-Voigt at 1e5
// draws x 4 clouds x 1000 channels.  The absorption kernel, one exponential per channel, gains 2 % and keeps one form.)
enum { EA_ALL = 0, EA_TAME = 1, EA_FIXUP = 2 };

template <int N, bool PV, bool SB, bool STAGED = false, int MODE = EA_ALL>
__global__ void __launch_bounds__(CHI_BLOCK, em_min_blocks(N, PV)) k_moments_em(ChanArgs a) {
  constexpr int NE = EmLayout<PV>::NE;
  constexpr bool SEL = MODE != EA_TAME;
  static_assert(!(STAGED && MODE == EA_FIXUP), "the scan form reads the channel data through L1");
  __shared__ __align__(16) double s_c[CHI_WARPS_PER_BLOCK][N][8];
  __shared__ double s_gb[SB ? 1 : CHI_MAX_BASELINE - CHI_BASE_REG][SB ? 1 : CHI_BLOCK];
  __shared__ double s_beta[CHI_WARPS_PER_BLOCK][CHI_MAX_BASELINE];
  __shared__ double s_T[CHI_EXP_TABLE];
  __shared__ __align__(128) double s_stage[STAGED ? CHI_WARPS_PER_BLOCK : 1][2][STAGED ? 128 : 2];
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
    const int nb = a.n_base, S = a.S;

    bool marked = false;
    if (lane < a.n_real) {
      const double* p = a.cc + (b * a.n_real + lane) * CHI_CC_STRIDE;
      const double f = p[CC_F], ts = p[CC_TS];
      double* d = s_c[warp][lane];
      d[0] = p[CC_C]; d[1] = -CHI_LOG2E * p[CC_KE]; d[2] = -CHI_LOG2E * p[CC_AGE]; d[3] = f * ts;
      d[4] = f; d[5] = ts; d[6] = -CHI_LOG2E * p[CC_ALE]; d[7] = p[CC_HE];
      marked = p[CC_CAREFUL] != 0.0;
    } else if (lane < N) {  // null cloud
      double* d = s_c[warp][lane];
      d[0] = 0.0; d[1] = -1.0; d[2] = 0.0; d[3] = 0.0; d[4] = 0.0; d[5] = 0.0; d[6] = 0.0; d[7] = 1.0;
    }
    if (MODE == EA_TAME && __any_sync(0xffffffffu, marked)) return;  // EA_FIXUP's draw
    if (lane < nb) s_beta[warp][lane] = a.coeff ? a.coeff[b * nb + lane] : 0.0;
    if (!SB)
      for (int i = CHI_BASE_REG; i < nb; ++i) s_gb[i - CHI_BASE_REG][threadIdx.x] = 0.0;
    __syncwarp();
    const double beta0 = nb > 0 ? s_beta[warp][0] : 0.0, beta1 = (!SB && nb > 1) ? s_beta[warp][1] : 0.0;
    const double* __restrict__ bas1 = a.basis + S;

    const int sl = a.sl ? a.sl[b] : 0;
    const double* __restrict__ ch = a.chan + (int64_t)sl * ((S + 31) >> 5) * 128 + lane;
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
    const unsigned stage_addr = (unsigned)__cvta_generic_to_shared(&s_stage[STAGED ? warp : 0][0][0]);
    const unsigned bar_addr = (unsigned)__cvta_generic_to_shared(&s_bar[STAGED ? warp : 0][0]);
    if (STAGED && lane == 0 && chunk * a.chunk_len < s_end)
      tma_fetch<1024>(ch - lane + (int64_t)((chunk * a.chunk_len) >> 5) * 128, stage_addr, bar_addr);
    // STAGED: the whole warp runs every group (lanes past a ragged end carry zero weight)
    for (int s = chunk * a.chunk_len + lane; (STAGED ? s - lane : s) < s_end; s += 32) {
      const double* __restrict__ cs = ch + (int64_t)(s >> 5) * 128;  // lane offset is in ch
      const int sx = STAGED ? min(s, s_end - 1) : s;
      const double* __restrict__ sg = cs;
      double v, p0;
      if (STAGED) {
        const int buf = it & 1;
        __syncwarp();  // every lane is done reading the other buffer
        if (lane == 0 && s - lane + 32 < s_end)
          tma_fetch<1024>(cs - lane + 128, stage_addr + (buf ^ 1) * 1024, bar_addr + (buf ^ 1) * 8);
        mbar_wait(bar_addr + buf * 8, (it >> 1) & 1);
        sg = &s_stage[warp][buf][lane];
        ++it;
        v = sg[0]; p0 = sg[32];
      } else {
        v = __ldg(cs); p0 = __ldg(cs + 32);
      }
      double p1 = 0.0, base = beta0 * p0;  // p_0 is 0 without a baseline
      if (!SB) {
        if (nb > 1) { p1 = bas1[s]; base = fma(beta1, p1, base); }
        for (int i = CHI_BASE_REG; i < nb; ++i) base = fma(s_beta[warp][i], a.basis[(int64_t)i * S + s], base);
      }

      // ---- forward, nearest cloud first ----
      double om[N], Pn[N], E[N], q[N];
      double P = 1.0, T = 0.0;
  #pragma unroll
      for (int n = 0; n < N; ++n) {
        double c0, nk, nAG, f, fTs, Ts;
        lds2(&sc[n][0], c0, nk);
        lds2(&sc[n][2], nAG, fTs);
        lds2(&sc[n][4], f, Ts);
        const double d = v - c0, d2 = d * d;
        E[n] = chi_exp2_prod<false>(nk, d2, ek);
        if (PV) {
          double x = nAG * E[n];                  // -tau log2(e)
          double nAL, h;
          lds2(&sc[n][6], nAL, h);
          q[n] = chi_rcp(d2 + h);
          x = fma(nAL, q[n], x);
          om[n] = 1.0 - chi_exp2<SEL, SEL>(x, ek);  // 1 - e^-tau
        } else {
          om[n] = 1.0 - chi_exp2_prod<SEL, SEL>(nAG, E[n], ek);
        }
        Pn[n] = P;
        const double u = om[n] * P;
        T = fma(fTs, u, T);                       // f Ts (1-e) x attenuation in front
        P = fma(-f, u, P);                        // P (1 - f + f e)
      }
      const double mu = (bg * P + T) - bg + base;  // ON - OFF, physics.py:362-365
      double mubar;
      if (G) {
        mubar = G[sx];
      } else {
        const double r = (STAGED ? sg[64] : __ldg(cs + 64)) - mu;
        mubar = (STAGED ? sg[96] : __ldg(cs + 96)) * r;  // d logp / d mu
        if (STAGED && s >= s_end) mubar = 0.0;    // padded lanes of a ragged end
        lp = fma(mubar, r, lp);                   // -0.5 applied once at the end
      }
      if (STAGED && s >= s_end) mubar = 0.0;
      gb0 = fma(mubar, p0, gb0);
      if (!SB) {
        if (nb > 1) gb1 = fma(mubar, p1, gb1);
        for (int i = CHI_BASE_REG; i < nb; ++i)
          s_gb[i - CHI_BASE_REG][threadIdx.x] =
              fma(mubar, a.basis[(int64_t)i * S + s], s_gb[i - CHI_BASE_REG][threadIdx.x]);
      }

      // ---- reverse, farthest cloud first; Q = brightness arriving behind cloud n ----
      double Q = bg;
  #pragma unroll
      for (int n = N - 1; n >= 0; --n) {
        double c0, nk, f, Ts;
        lds2(&sc[n][0], c0, nk);
        lds2(&sc[n][4], f, Ts);
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
        Q = fma(om[n], f * Sn, Q);                // (1 - f om) Q + f Ts om
      }
    }

    double* out = a.mom + item * a.mom_stride;
    if (!SB)
      for (int i = CHI_BASE_REG; i < nb; ++i) {
        double g = warp_sum(s_gb[i - CHI_BASE_REG][threadIdx.x]);
        if (lane == 0) out[1 + i] = g;
      }
    // logp, the two register-held baseline gradients and the N * NE moments: one folded reduction
    typedef WarpFold<3 + N * NE> Fold;
    double v[3 + N * NE];
    v[0] = lp; v[1] = gb0; v[2] = gb1;
  #pragma unroll
    for (int n = 0; n < N; ++n)
  #pragma unroll
      for (int j = 0; j < NE; ++j) v[3 + n * NE + j] = acc[n][j];
    Fold::run(v, lane);
  #pragma unroll
    for (int i = 0; i < Fold::OUT; ++i) {
      const int idx = Fold::index(i, lane);
      if (idx >= 3) out[1 + nb + (idx - 3)] = v[i];
      else if (idx == 0) out[0] = -0.5 * v[i];
      else if (idx > 0 && idx <= nb) out[idx] = v[i];
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
        const int64_t bb = b0 + (__ffs(todo) - 1);
        for (int chunk = 0; chunk < a.n_chunks; ++chunk) {
          one_item(bb * a.n_chunks + chunk);
          __syncwarp();  // the warp's shared constants are rewritten by the next item
        }
      }
    }
  }
}

// ---------------------------------------------------------------------------
// k_moments_abs: absorption spectrum 1 - exp(-sum_n tau)
//   absorption_model.py:88-91, emission_absorption_model.py:84-99 + physical equivalents
//   shared per cloud: c, -k', -AG', -AL', h   (' = times log2(e), for chi_exp2)
// ---------------------------------------------------------------------------
template <int N, bool PV, bool SB, bool STAGED = false>
__global__ void __launch_bounds__(CHI_BLOCK, ab_min_blocks(N, PV)) k_moments_abs(ChanArgs a) {
  constexpr int NA = AbLayout<PV>::NA;
  __shared__ __align__(16) double s_c[CHI_WARPS_PER_BLOCK][N][6];
  __shared__ double s_gb[SB ? 1 : CHI_MAX_BASELINE - CHI_BASE_REG][SB ? 1 : CHI_BLOCK];
  __shared__ double s_beta[CHI_WARPS_PER_BLOCK][CHI_MAX_BASELINE];
  __shared__ double s_T[CHI_EXP_TABLE];
  __shared__ __align__(128) double s_stage[STAGED ? CHI_WARPS_PER_BLOCK : 1][2][STAGED ? 128 : 2];
  __shared__ __align__(8) unsigned long long s_bar[STAGED ? CHI_WARPS_PER_BLOCK : 1][2];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (STAGED && lane == 0) {
    mbar_init((unsigned)__cvta_generic_to_shared(&s_bar[warp][0]));
    mbar_init((unsigned)__cvta_generic_to_shared(&s_bar[warp][1]));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  fill_exp_table(s_T);  // contains the block barrier

  const int64_t item = (int64_t)blockIdx.x * CHI_WARPS_PER_BLOCK + warp;
  if (item >= a.B * a.n_chunks) return;
  const int64_t b = item / a.n_chunks;
  const int chunk = (int)(item - b * a.n_chunks);
  const int nb = a.n_base, S = a.S;

  if (lane < a.n_real) {
    const double* p = a.cc + (b * a.n_real + lane) * CHI_CC_STRIDE;
    double* d = s_c[warp][lane];
    d[0] = p[CC_C]; d[1] = -CHI_LOG2E * p[CC_KA]; d[2] = -CHI_LOG2E * p[CC_AGA];
    d[3] = -CHI_LOG2E * p[CC_ALA]; d[4] = p[CC_HA];
    d[5] = 0.0;
  } else if (lane < N) {  // null cloud
    double* d = s_c[warp][lane];
    d[0] = 0.0; d[1] = -1.0; d[2] = 0.0; d[3] = 0.0; d[4] = 1.0; d[5] = 0.0;
  }
  if (lane < nb) s_beta[warp][lane] = a.coeff ? a.coeff[b * nb + lane] : 0.0;
  if (!SB)
    for (int i = CHI_BASE_REG; i < nb; ++i) s_gb[i - CHI_BASE_REG][threadIdx.x] = 0.0;
  __syncwarp();
  const double beta0 = nb > 0 ? s_beta[warp][0] : 0.0, beta1 = (!SB && nb > 1) ? s_beta[warp][1] : 0.0;
  const double* __restrict__ bas1 = a.basis + S;

  const int sl = a.sl ? a.sl[b] : 0;
  const double* __restrict__ ch = a.chan + (int64_t)sl * ((S + 31) >> 5) * 128 + lane;
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
  const unsigned stage_addr = (unsigned)__cvta_generic_to_shared(&s_stage[STAGED ? warp : 0][0][0]);
  const unsigned bar_addr = (unsigned)__cvta_generic_to_shared(&s_bar[STAGED ? warp : 0][0]);
  int it = 0;
  if (STAGED && lane == 0 && chunk * a.chunk_len < s_end)
    tma_fetch<1024>(ch - lane + (int64_t)((chunk * a.chunk_len) >> 5) * 128, stage_addr, bar_addr);
  // STAGED: the whole warp runs every group (lanes past a ragged end carry zero weight)
  for (int s = chunk * a.chunk_len + lane; (STAGED ? s - lane : s) < s_end; s += 32) {
    const double* __restrict__ cs = ch + (int64_t)(s >> 5) * 128;  // lane offset is in ch
    const int sx = STAGED ? min(s, s_end - 1) : s;
    const double* __restrict__ sg = cs;
    double v, p0;
    if (STAGED) {
      const int buf = it & 1;
      __syncwarp();  // every lane is done reading the other buffer
      if (lane == 0 && s - lane + 32 < s_end)
        tma_fetch<1024>(cs - lane + 128, stage_addr + (buf ^ 1) * 1024, bar_addr + (buf ^ 1) * 8);
      mbar_wait(bar_addr + buf * 8, (it >> 1) & 1);
      sg = &s_stage[warp][buf][lane];
      ++it;
      v = sg[0]; p0 = sg[32];
    } else {
      v = __ldg(cs); p0 = __ldg(cs + 32);
    }
    double p1 = 0.0, base = beta0 * p0;  // p_0 is 0 without a baseline
    if (!SB) {
      if (nb > 1) { p1 = bas1[s]; base = fma(beta1, p1, base); }
      for (int i = CHI_BASE_REG; i < nb; ++i) base = fma(s_beta[warp][i], a.basis[(int64_t)i * S + s], base);
    }

    double E[N], d[N], q[N];
    double xs = 0.0;  // -sum_n tau_n
#pragma unroll
    for (int n = 0; n < N; ++n) {
      double c0, nk, nAG, nAL;
      lds2(&sc[n][0], c0, nk);
      lds2(&sc[n][2], nAG, nAL);
      d[n] = v - c0;
      const double d2 = d[n] * d[n];
      E[n] = chi_exp2_prod<false>(nk, d2, ek);
      xs = fma(nAG, E[n], xs);
      if (PV) {
        q[n] = chi_rcp(d2 + sc[n][4]);
        xs = fma(nAL, q[n], xs);
      }
    }
    const double ea = chi_exp2(xs, ek);
    const double mu = (1.0 - ea) + base;
    double mubar;
    if (G) {
      mubar = G[sx];
    } else {
      const double r = (STAGED ? sg[64] : __ldg(cs + 64)) - mu;
      mubar = (STAGED ? sg[96] : __ldg(cs + 96)) * r;
      if (STAGED && s >= s_end) mubar = 0.0;    // padded lanes of a ragged end
      lp = fma(mubar, r, lp);
    }
    if (STAGED && s >= s_end) mubar = 0.0;
    gb0 = fma(mubar, p0, gb0);
    if (!SB) {
      if (nb > 1) gb1 = fma(mubar, p1, gb1);
      for (int i = CHI_BASE_REG; i < nb; ++i)
        s_gb[i - CHI_BASE_REG][threadIdx.x] =
            fma(mubar, a.basis[(int64_t)i * S + s], s_gb[i - CHI_BASE_REG][threadIdx.x]);
    }

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
  if (!SB)
    for (int i = CHI_BASE_REG; i < nb; ++i) {
      double g = warp_sum(s_gb[i - CHI_BASE_REG][threadIdx.x]);
      if (lane == 0) out[1 + i] = g;
    }
  // logp, the two register-held baseline gradients and the N * NA moments: one folded reduction
  typedef WarpFold<3 + N * NA> Fold;
  double v[3 + N * NA];
  v[0] = lp; v[1] = gb0; v[2] = gb1;
#pragma unroll
  for (int n = 0; n < N; ++n)
#pragma unroll
    for (int j = 0; j < NA; ++j) v[3 + n * NA + j] = acc[n][j];
  Fold::run(v, lane);
#pragma unroll
  for (int i = 0; i < Fold::OUT; ++i) {
    const int idx = Fold::index(i, lane);
    if (idx >= 3) out[1 + nb + (idx - 3)] = v[i];
    else if (idx == 0) out[0] = -0.5 * v[i];
    else if (idx > 0 && idx <= nb) out[idx] = v[i];
  }
}

}  // namespace chi
