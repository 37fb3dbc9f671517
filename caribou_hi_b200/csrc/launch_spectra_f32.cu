// k_spectra32_t<N, PV, MODE>: predicted spectra in single precision -- the optional fp32 path of
// the north star, for the one output where fp32 can meet its 1e-5 bar (posterior / prior
// predictive spectra, BASELINE configs[4]; logp and gradients stay fp64).
// emission: physics.py:294-309 + physics.py:345-365; absorption: absorption_model.py:88-91.
//
// The parameter maps (k_cloud_map) stay fp64; this kernel converts the per-(draw, cloud)
// constants to float once per block and runs the O(B N S) channel loop on the FP32 pipe and the
// SFU (MUFU.EX2 / MUFU.RCP).  Three places where a naive fp32 translation misses 1e-5, and what
// is done instead:
//   * d = v - c: rounding v (|v| ~ 100 km/s) and c to float costs ~5e-6 km/s, which a narrow
//     line (W ~ one channel) turns into 1e-4 of its peak.  v and c are carried as (hi, lo) float
//     pairs, d = (v_hi - c_hi) + (v_lo - c_lo): the first difference is correctly rounded
//     relative to d itself.
//   * 1 - exp(-tau) for optically thin warm gas (tau ~ 1e-3, Ts ~ 1e4 K): the cancellation
//     leaves 6e-8 * Ts absolute.  For -tau' > -1/8 a cubic-times-x polynomial of 1 - 2^x is
//     selected (branch-free) instead of 1 - ex2(x).
//   * bg * P - bg with P = prod(1 - f + f e^-tau) ~ 1: the kernel accumulates R = 1 - P beside P.
// Adjacent channel pairs per thread (float4 axis loads, float2 stores).
#include "launch.cuh"

namespace chi {

__device__ __forceinline__ float ex2f(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float rcpf(float x) {
  float y;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
// 1 - 2^x to ~1.5e-6 relative: cubic near 0 (no cancellation), 1 - ex2 elsewhere
__device__ __forceinline__ float one_minus_exp2(float x) {
  const float direct = 1.0f - ex2f(x);
  // -(2^x - 1) = -x (c1 + x (c2 + x (c3 + x c4))), c_k = ln2^k / k!; |x| <= 1/8: next term 5e-7
  float p = fmaf(x, 9.6181291e-3f, 5.5504109e-2f);
  p = fmaf(x, p, 2.4022651e-1f);
  p = fmaf(x, p, 6.9314718e-1f);
  return fabsf(x) < 0.125f ? -x * p : direct;  // |x|: a negative optical depth (physics-level callers) stays valid
}

#ifndef CHI_S32_MINB
#define CHI_S32_MINB 5
#endif
#ifndef CHI_S32_CPT
#define CHI_S32_CPT 4
#endif

// MODE 0: emission only, 1: absorption only, 2: both spectra of a shared axis in one pass.
// A thread carries CPT channels (CPT/2 adjacent pairs, 512 channels apart) through the cloud loop,
// so the constants of a cloud are loaded once per CPT channels.
template <int N, bool PV, int MODE>
__global__ void __launch_bounds__(256, (N <= 8 ? CHI_S32_MINB : 3)) k_spectra32_t(Spec32Args a) {
  constexpr bool EM = MODE != 1;
  constexpr int CPT = CHI_S32_CPT, NP = CPT / 2;
  // per cloud: [c_hi, c_lo, -k', -AG_e' (-AG_a' in MODE 1)] [f*Ts, f, -AG_a', h] [-AL_e' (-AL_a' in MODE 1), -AL_a', -, -]
  __shared__ __align__(16) float s_c[N][12];
  __shared__ double s_beta[2][CHI_MAX_BASELINE];
  const int64_t b = blockIdx.x;
  if (threadIdx.x < N) {
    float* d = s_c[threadIdx.x];
    if (threadIdx.x >= a.N) {  // null cloud of a padded instantiation
      for (int i = 0; i < 12; ++i) d[i] = 0.0f;
      d[2] = -1.0f; d[7] = 1.0f;
    } else {
      const double* p = a.cc + (b * a.N + threadIdx.x) * CHI_CC_STRIDE;
      const double c = p[CC_C];
      const float ch = (float)c;
      d[0] = ch; d[1] = (float)(c - (double)ch);
      if (MODE == 1) {
        d[2] = (float)(-CHI_LOG2E * p[CC_KA]); d[3] = (float)(-CHI_LOG2E * p[CC_AGA]);
        d[4] = 0.0f; d[5] = 0.0f; d[6] = 0.0f; d[7] = (float)p[CC_HA];
        d[8] = (float)(-CHI_LOG2E * p[CC_ALA]); d[9] = 0.0f;
      } else {
        d[2] = (float)(-CHI_LOG2E * p[CC_KE]); d[3] = (float)(-CHI_LOG2E * p[CC_AGE]);
        d[4] = (float)(p[CC_F] * p[CC_TS]); d[5] = (float)p[CC_F];
        d[6] = (float)(-CHI_LOG2E * p[CC_AGA]); d[7] = (float)p[CC_HE];
        d[8] = (float)(-CHI_LOG2E * p[CC_ALE]); d[9] = (float)(-CHI_LOG2E * p[CC_ALA]);
      }
      d[10] = d[11] = 0.0f;
    }
  }
  if (threadIdx.x < a.n_base) s_beta[0][threadIdx.x] = a.coeff ? a.coeff[b * a.n_base + threadIdx.x] : 0.0;
  if (MODE == 2 && threadIdx.x < a.n_base_a)
    s_beta[1][threadIdx.x] = a.coeff_a ? a.coeff_a[b * a.n_base_a + threadIdx.x] : 0.0;
  __syncthreads();
  const int S = a.S;
  const bool even = (S & 1) == 0;
  float* __restrict__ out0 = a.mu + b * S;      // emission (MODE 0, 2) or absorption (MODE 1)
  float* __restrict__ out1 = a.mu_a + b * S;    // absorption of MODE 2
  const float4* __restrict__ vax = reinterpret_cast<const float4*>(a.vax32);  // padded to an even channel count
  const float nbg = -a.bg;
  for (int s0 = 2 * threadIdx.x; s0 < S; s0 += 512 * NP) {
    float vh[CPT], vl[CPT];
#pragma unroll
    for (int p = 0; p < NP; ++p) {
      const int s = s0 + 512 * p;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (s < S) v = vax[s >> 1];  // (hi, lo) of channels s and s + 1
      vh[2 * p] = v.x; vl[2 * p] = v.y; vh[2 * p + 1] = v.z; vl[2 * p + 1] = v.w;
    }
    float R[CPT], P[CPT], T[CPT], xs[CPT];
#pragma unroll
    for (int j = 0; j < CPT; ++j) { R[j] = 0.0f; P[j] = 1.0f; T[j] = 0.0f; xs[j] = 0.0f; }
#pragma unroll
    for (int n = 0; n < N; ++n) {
      const float4 q0 = *reinterpret_cast<const float4*>(&s_c[n][0]);
      const float4 q1 = *reinterpret_cast<const float4*>(&s_c[n][4]);
      float2 q2 = make_float2(0.0f, 0.0f);
      if (PV) q2 = *reinterpret_cast<const float2*>(&s_c[n][8]);
#pragma unroll
      for (int j = 0; j < CPT; ++j) {
        const float d = (vh[j] - q0.x) + (vl[j] - q0.y);
        const float d2 = d * d;
        const float E = ex2f(q0.z * d2);
        float x = q0.w * E;  // -tau log2(e) of the first dataset
        if (MODE == 2) xs[j] = fmaf(q1.z, E, xs[j]);
        if (PV) {
          const float q = rcpf(d2 + q1.w);
          x = fmaf(q2.x, q, x);
          if (MODE == 2) xs[j] = fmaf(q2.y, q, xs[j]);
        }
        if (EM) {
          const float u = one_minus_exp2(x) * P[j];
          T[j] = fmaf(q1.x, u, T[j]);
          R[j] = fmaf(q1.y, u, R[j]);   // 1 - P without the cancellation
          P[j] = fmaf(-q1.y, u, P[j]);
        } else {
          xs[j] += x;
        }
      }
    }
#pragma unroll
    for (int p = 0; p < NP; ++p) {
      const int s = s0 + 512 * p;
      if (s >= S) break;
      const bool two = s + 1 < S;
      float r0[2], r1[2] = {0.0f, 0.0f};
#pragma unroll
      for (int jj = 0; jj < 2; ++jj) {
        const int j = 2 * p + jj;
        const int sj = two ? s + jj : s;
        float base0 = 0.0f, base1 = 0.0f;
        if (a.offset || a.n_base) {
          double t = a.offset ? a.offset[sj] : 0.0;
#pragma unroll 1
          for (int i = 0; i < a.n_base; ++i) t = fma(s_beta[0][i], a.basis[(int64_t)i * S + sj], t);
          base0 = (float)t;
        }
        if (MODE == 2 && (a.offset_a || a.n_base_a)) {
          double t = a.offset_a ? a.offset_a[sj] : 0.0;
#pragma unroll 1
          for (int i = 0; i < a.n_base_a; ++i) t = fma(s_beta[1][i], a.basis_a[(int64_t)i * S + sj], t);
          base1 = (float)t;
        }
        if (EM) r0[jj] = fmaf(nbg, R[j], T[j]) + base0;
        else r0[jj] = one_minus_exp2(xs[j]) + base0;
        if (MODE == 2) r1[jj] = one_minus_exp2(xs[j]) + base1;
      }
      if (even) {
        *reinterpret_cast<float2*>(out0 + s) = make_float2(r0[0], r0[1]);
        if (MODE == 2) *reinterpret_cast<float2*>(out1 + s) = make_float2(r1[0], r1[1]);
      } else {
        out0[s] = r0[0];
        if (two) out0[s + 1] = r0[1];
        if (MODE == 2) {
          out1[s] = r1[0];
          if (two) out1[s + 1] = r1[1];
        }
      }
    }
  }
}

template <int N, bool PV>
static void run32(const Spec32Args& a, cudaStream_t st) {
  if (a.mode == 2) k_spectra32_t<N, PV, 2><<<(unsigned)a.B, 256, 0, st>>>(a);
  else if (a.mode == 1) k_spectra32_t<N, PV, 1><<<(unsigned)a.B, 256, 0, st>>>(a);
  else k_spectra32_t<N, PV, 0><<<(unsigned)a.B, 256, 0, st>>>(a);
}

int launch_spectra32(int N, bool pv, const Spec32Args& a, cudaStream_t st) {
  N = padded_clouds(N);
  switch (N * 2 + (pv ? 1 : 0)) {
#define CASE(n)                             \
  case (n)*2: run32<n, false>(a, st); return 0; \
  case (n)*2 + 1: run32<n, true>(a, st); return 0;
    CASE(1) CASE(2) CASE(3) CASE(4) CASE(5) CASE(6) CASE(7) CASE(8) CASE(12) CASE(16)
#undef CASE
  }
  return -1;
}
}  // namespace chi
