// k_spectra_t<N, PV, EMISSION>: predicted spectra only (posterior / prior predictive,
// BASELINE configs[4]) -- the forward half of moments.cuh with the cloud loop unrolled (the N
// independent exp chains give the instruction-level parallelism), the table-driven branch-free
// exp and the Newton reciprocal; one block per draw, threads walk channels, 8-byte coalesced
// stores of mu.  emission: physics.py:294-309 + physics.py:345-365; absorption:
// absorption_model.py:88-91.
#include <type_traits>

#include "launch.cuh"

namespace chi {

template <int N, bool PV, bool EMISSION>
__global__ void __launch_bounds__(256) k_spectra_t(SpecArgs a) {
  __shared__ __align__(16) double s_c[N][8];
  __shared__ double s_beta[CHI_MAX_BASELINE];
  __shared__ double s_T[CHI_EXP_TABLE];
  const int64_t b = blockIdx.x;
  if (threadIdx.x >= a.N && threadIdx.x < N) {  // null cloud of a padded instantiation
    double* d = s_c[threadIdx.x];
    d[0] = 0.0; d[1] = -1.0; d[2] = 0.0; d[3] = 0.0; d[4] = 1.0; d[5] = d[6] = d[7] = 0.0;
  }
  bool marked = false;
  if (threadIdx.x < a.N) {
    const double* p = a.cc + (b * a.N + threadIdx.x) * CHI_CC_STRIDE;
    double* d = s_c[threadIdx.x];
    marked = p[CC_CAREFUL] != 0.0;
    d[0] = p[CC_C];
    if (EMISSION) {
      d[1] = -CHI_LOG2E * p[CC_KE]; d[2] = -CHI_LOG2E * p[CC_AGE]; d[3] = -CHI_LOG2E * p[CC_ALE]; d[4] = p[CC_HE];
      d[5] = p[CC_F]; d[6] = p[CC_TS]; d[7] = p[CC_F] * p[CC_TS];
    } else {
      d[1] = -CHI_LOG2E * p[CC_KA]; d[2] = -CHI_LOG2E * p[CC_AGA]; d[3] = -CHI_LOG2E * p[CC_ALA]; d[4] = p[CC_HA];
      d[5] = d[6] = d[7] = 0.0;
    }
  }
  if (threadIdx.x < a.n_base) s_beta[threadIdx.x] = a.coeff ? a.coeff[b * a.n_base + threadIdx.x] : 0.0;
  fill_exp_table(s_T);  // contains the block barrier
  const ExpCtx ek = CHI_LOAD_EXPCTX(s_T);
  const int S = a.S;
  double* __restrict__ out = a.mu + b * S;
  // The range selects of exp(-tau) (overflow: a negative depth; underflow: a huge one) are only compiled into the
  // copy of the loop that draws marked by the cloud map take (cc[..][CC_CAREFUL], kernels.cuh); the choice is
  // uniform over the block.  14 of ~70 instructions per (channel, cloud) less for everybody else.
  const bool careful = __syncthreads_or(marked) != 0;
  auto channels = [&](auto tag) {
  constexpr bool SEL = decltype(tag)::value;
  for (int s = threadIdx.x; s < S; s += blockDim.x) {
    const double v = a.vax[s];
    double base = a.offset ? a.offset[s] : 0.0;
    for (int i = 0; i < a.n_base; ++i) base = fma(s_beta[i], a.basis[(int64_t)i * S + s], base);
    double E[N], x[N];
#pragma unroll
    for (int n = 0; n < N; ++n) {
      double c0, nk, nAG, nAL;
      lds2(&s_c[n][0], c0, nk);
      lds2(&s_c[n][2], nAG, nAL);
      const double d = v - c0, d2 = d * d;
      E[n] = chi_exp2_prod<false>(nk, d2, ek);
      x[n] = nAG * E[n];  // -tau log2(e)
      if (PV) x[n] = fma(nAL, chi_rcp(d2 + s_c[n][4]), x[n]);
    }
    double r;
    if (EMISSION) {
      double P = 1.0, T = 0.0;
#pragma unroll
      for (int n = 0; n < N; ++n) {
        const double u = (1.0 - chi_exp2<SEL, SEL>(x[n], ek)) * P;
        T = fma(s_c[n][7], u, T);
        P = fma(-s_c[n][5], u, P);
      }
      r = (a.bg * P + T) - a.bg + base;
    } else {
      double xs = 0.0;
#pragma unroll
      for (int n = 0; n < N; ++n) xs += x[n];
      r = (1.0 - chi_exp2<SEL, SEL>(xs, ek)) + base;
    }
    out[s] = r;
  }
  };
  if (careful) channels(std::true_type{});
  else channels(std::false_type{});
}

// Both spectra of a shared axis in one pass: exp(-k d^2) of a cloud is evaluated once and feeds
// the transfer (emission) and the summed optical depth (absorption), like k_moments_ea.
template <int N, bool PV>
__global__ void __launch_bounds__(256) k_spectra_ea_t(SpecArgs a) {
  // per cloud: [c, -k'] [-AG_e', f*Ts] [f, -AG_a'] [-AL_e', -AL_a'] [h, -]
  __shared__ __align__(16) double s_c[N][10];
  __shared__ double s_beta[2][CHI_MAX_BASELINE];
  __shared__ double s_T[CHI_EXP_TABLE];
  const int64_t b = blockIdx.x;
  if (threadIdx.x >= a.N && threadIdx.x < N) {  // null cloud of a padded instantiation
    double* d = s_c[threadIdx.x];
    for (int i = 0; i < 10; ++i) d[i] = 0.0;
    d[1] = -1.0; d[8] = 1.0;
  }
  bool marked = false;
  if (threadIdx.x < a.N) {
    const double* p = a.cc + (b * a.N + threadIdx.x) * CHI_CC_STRIDE;
    double* d = s_c[threadIdx.x];
    marked = p[CC_CAREFUL] != 0.0;
    d[0] = p[CC_C]; d[1] = -CHI_LOG2E * p[CC_KE]; d[2] = -CHI_LOG2E * p[CC_AGE]; d[3] = p[CC_F] * p[CC_TS];
    d[4] = p[CC_F]; d[5] = -CHI_LOG2E * p[CC_AGA]; d[6] = -CHI_LOG2E * p[CC_ALE]; d[7] = -CHI_LOG2E * p[CC_ALA];
    d[8] = p[CC_HE]; d[9] = 0.0;
  }
  if (threadIdx.x < a.n_base) s_beta[0][threadIdx.x] = a.coeff ? a.coeff[b * a.n_base + threadIdx.x] : 0.0;
  if (threadIdx.x < a.n_base_a) s_beta[1][threadIdx.x] = a.coeff_a ? a.coeff_a[b * a.n_base_a + threadIdx.x] : 0.0;
  fill_exp_table(s_T);  // contains the block barrier
  const ExpCtx ek = CHI_LOAD_EXPCTX(s_T);
  const int S = a.S;
  double* __restrict__ out_e = a.mu + b * S;
  double* __restrict__ out_a = a.mu_a + b * S;
  const bool careful = __syncthreads_or(marked) != 0;  // as in k_spectra_t
  auto channels = [&](auto tag) {
  constexpr bool SEL = decltype(tag)::value;
  for (int s = threadIdx.x; s < S; s += blockDim.x) {
    const double v = a.vax[s];
    double base_e = a.offset ? a.offset[s] : 0.0, base_a = a.offset_a ? a.offset_a[s] : 0.0;
    for (int i = 0; i < a.n_base; ++i) base_e = fma(s_beta[0][i], a.basis[(int64_t)i * S + s], base_e);
    for (int i = 0; i < a.n_base_a; ++i) base_a = fma(s_beta[1][i], a.basis_a[(int64_t)i * S + s], base_a);
    double P = 1.0, T = 0.0, xs = 0.0;
#pragma unroll
    for (int n = 0; n < N; ++n) {
      double c0, nk, nAGe, fTs, f, nAGa;
      lds2(&s_c[n][0], c0, nk);
      lds2(&s_c[n][2], nAGe, fTs);
      lds2(&s_c[n][4], f, nAGa);
      const double d = v - c0, d2 = d * d;
      const double E = chi_exp2_prod<false>(nk, d2, ek);
      double x = nAGe * E;
      xs = fma(nAGa, E, xs);
      if (PV) {
        double nALe, nALa;
        lds2(&s_c[n][6], nALe, nALa);
        const double q = chi_rcp(d2 + s_c[n][8]);
        x = fma(nALe, q, x);
        xs = fma(nALa, q, xs);
      }
      const double u = (1.0 - chi_exp2<SEL, SEL>(x, ek)) * P;
      T = fma(fTs, u, T);
      P = fma(-f, u, P);
    }
    out_e[s] = (a.bg * P + T) - a.bg + base_e;
    out_a[s] = (1.0 - chi_exp2<SEL, SEL>(xs, ek)) + base_a;
  }
  };
  if (careful) channels(std::true_type{});
  else channels(std::false_type{});
}

template <int N, bool PV>
static void run_ea(const SpecArgs& a, cudaStream_t st) {
  k_spectra_ea_t<N, PV><<<(unsigned)a.B, 256, 0, st>>>(a);
}

int launch_spectra_ea(int N, bool pv, const SpecArgs& a, cudaStream_t st) {
  N = padded_clouds(N);
  switch (N * 2 + (pv ? 1 : 0)) {
#define CASE(n)                              \
  case (n)*2: run_ea<n, false>(a, st); return 0; \
  case (n)*2 + 1: run_ea<n, true>(a, st); return 0;
    CASE(1) CASE(2) CASE(3) CASE(4) CASE(5) CASE(6) CASE(7) CASE(8) CASE(12) CASE(16)
#undef CASE
  }
  return -1;
}

template <int N, bool PV>
static void run(const SpecArgs& a, cudaStream_t st) {
  if (a.emission) k_spectra_t<N, PV, true><<<(unsigned)a.B, 256, 0, st>>>(a);
  else k_spectra_t<N, PV, false><<<(unsigned)a.B, 256, 0, st>>>(a);
}

int launch_spectra(int N, bool pv, const SpecArgs& a, cudaStream_t st) {
  N = padded_clouds(N);
  switch (N * 2 + (pv ? 1 : 0)) {
#define CASE(n)                           \
  case (n)*2: run<n, false>(a, st); return 0; \
  case (n)*2 + 1: run<n, true>(a, st); return 0;
    CASE(1) CASE(2) CASE(3) CASE(4) CASE(5) CASE(6) CASE(7) CASE(8) CASE(12) CASE(16)
#undef CASE
  }
  return -1;
}
}  // namespace chi
