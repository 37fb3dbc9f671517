// Forward-mode dual numbers for the O(N) per-cloud parameter maps.
//
// The O(N*S) part of the gradient (profile, radiative transfer, logp) is a
// hand-derived reverse pass (see moments.cuh).  The O(N) per-cloud maps
// (free RVs -> spin temperature -> profile constants) have at most CHI_NSLOT inputs
// per cloud, so their Jacobian is taken by running the same templated code on
// Dual<K>: value + K tangents.  The same source instantiated with `double` gives
// the value-only prologue, so forward and derivative code cannot drift apart.
#pragma once
#include <cuda_runtime.h>
#include <math.h>

namespace chi {

template <int K>
struct Dual {
  double v;
  double d[K];
  __host__ __device__ Dual() {}
  __host__ __device__ Dual(double x) : v(x) {
#pragma unroll
    for (int i = 0; i < K; ++i) d[i] = 0.0;
  }
  // independent variable number `slot`
  __host__ __device__ static Dual variable(double x, int slot) {
    Dual r(x);
#pragma unroll
    for (int i = 0; i < K; ++i) r.d[i] = (i == slot) ? 1.0 : 0.0;
    return r;
  }
};

#define CHI_HD __host__ __device__ __forceinline__

template <int K> CHI_HD Dual<K> operator+(const Dual<K>& a, const Dual<K>& b) {
  Dual<K> r; r.v = a.v + b.v;
#pragma unroll
  for (int i = 0; i < K; ++i) r.d[i] = a.d[i] + b.d[i];
  return r;
}
template <int K> CHI_HD Dual<K> operator+(const Dual<K>& a, double b) { Dual<K> r = a; r.v += b; return r; }
template <int K> CHI_HD Dual<K> operator+(double a, const Dual<K>& b) { return b + a; }
template <int K> CHI_HD Dual<K> operator-(const Dual<K>& a) {
  Dual<K> r; r.v = -a.v;
#pragma unroll
  for (int i = 0; i < K; ++i) r.d[i] = -a.d[i];
  return r;
}
template <int K> CHI_HD Dual<K> operator-(const Dual<K>& a, const Dual<K>& b) {
  Dual<K> r; r.v = a.v - b.v;
#pragma unroll
  for (int i = 0; i < K; ++i) r.d[i] = a.d[i] - b.d[i];
  return r;
}
template <int K> CHI_HD Dual<K> operator-(const Dual<K>& a, double b) { Dual<K> r = a; r.v -= b; return r; }
template <int K> CHI_HD Dual<K> operator-(double a, const Dual<K>& b) { Dual<K> r = -b; r.v += a; return r; }
template <int K> CHI_HD Dual<K> operator*(const Dual<K>& a, const Dual<K>& b) {
  Dual<K> r; r.v = a.v * b.v;
#pragma unroll
  for (int i = 0; i < K; ++i) r.d[i] = a.d[i] * b.v + a.v * b.d[i];
  return r;
}
template <int K> CHI_HD Dual<K> operator*(const Dual<K>& a, double b) {
  Dual<K> r; r.v = a.v * b;
#pragma unroll
  for (int i = 0; i < K; ++i) r.d[i] = a.d[i] * b;
  return r;
}
template <int K> CHI_HD Dual<K> operator*(double a, const Dual<K>& b) { return b * a; }
template <int K> CHI_HD Dual<K> operator/(const Dual<K>& a, const Dual<K>& b) {
  Dual<K> r; r.v = a.v / b.v;
  const double ib = 1.0 / b.v;
#pragma unroll
  for (int i = 0; i < K; ++i) r.d[i] = (a.d[i] - r.v * b.d[i]) * ib;
  return r;
}
template <int K> CHI_HD Dual<K> operator/(const Dual<K>& a, double b) {
  Dual<K> r; r.v = a.v / b;
#pragma unroll
  for (int i = 0; i < K; ++i) r.d[i] = a.d[i] / b;
  return r;
}
template <int K> CHI_HD Dual<K> operator/(double a, const Dual<K>& b) {
  Dual<K> r; r.v = a / b.v;
  const double s = -r.v / b.v;
#pragma unroll
  for (int i = 0; i < K; ++i) r.d[i] = s * b.d[i];
  return r;
}

// chain rule helper: f(a) with derivative df
template <int K> CHI_HD Dual<K> chain(const Dual<K>& a, double f, double df) {
  Dual<K> r; r.v = f;
#pragma unroll
  for (int i = 0; i < K; ++i) r.d[i] = df * a.d[i];
  return r;
}

// ---- math on double and Dual with one spelling ---------------------------------
CHI_HD double m_value(double a) { return a; }
template <int K> CHI_HD double m_value(const Dual<K>& a) { return a.v; }

CHI_HD double m_sqrt(double a) { return sqrt(a); }
template <int K> CHI_HD Dual<K> m_sqrt(const Dual<K>& a) { double s = sqrt(a.v); return chain(a, s, 0.5 / s); }

CHI_HD double m_log(double a) { return log(a); }
template <int K> CHI_HD Dual<K> m_log(const Dual<K>& a) { return chain(a, log(a.v), 1.0 / a.v); }

#define CHI_LN10 2.302585092994045684
CHI_HD double m_log10(double a) { return log10(a); }
template <int K> CHI_HD Dual<K> m_log10(const Dual<K>& a) { return chain(a, log10(a.v), 1.0 / (a.v * CHI_LN10)); }

CHI_HD double m_exp(double a) { return exp(a); }
template <int K> CHI_HD Dual<K> m_exp(const Dual<K>& a) { double e = exp(a.v); return chain(a, e, e); }

// 10**a  (the reference writes 10.0**x; PyTensor evaluates pow(10, x))
CHI_HD double m_exp10(double a) { return exp10(a); }
template <int K> CHI_HD Dual<K> m_exp10(const Dual<K>& a) { double e = exp10(a.v); return chain(a, e, e * CHI_LN10); }

// a**c with constant exponent c > 0 and a**b with both variable, a >= 0.  Evaluated as
// exp(c log a) (a sqrt(a) for c = 1.5): libm's pow is several times longer and these sit on the
// latency-critical O(N) chain of small batches; the relative difference to pow is ~|c log a| ulp
// (< 1e-14 here), special values (0, inf, NaN, a < 0) come out the same for c > 0.
CHI_HD double pow_pos(double a, double c) { return c == 1.5 ? a * sqrt(a) : exp(c * log(a)); }
CHI_HD double m_powc(double a, double c) { return pow_pos(a, c); }
template <int K> CHI_HD Dual<K> m_powc(const Dual<K>& a, double c) {
  double p = pow_pos(a.v, c);
  return chain(a, p, c * p / a.v);
}
CHI_HD double m_pow(double a, double b) { return exp(b * log(a)); }
template <int K> CHI_HD Dual<K> m_pow(const Dual<K>& a, const Dual<K>& b) {
  Dual<K> r; r.v = exp(b.v * log(a.v));
  const double da = b.v * r.v / a.v, db = r.v * log(a.v);
#pragma unroll
  for (int i = 0; i < K; ++i) r.d[i] = da * a.d[i] + db * b.d[i];
  return r;
}

// switch(cond, a, b): gradient flows through the selected branch only
template <class T> CHI_HD T m_select(bool c, const T& a, const T& b) { return c ? a : b; }

// clip(x, lo, hi): PyTensor's clip passes gradient where lo <= x <= hi, else 0
CHI_HD double m_clip(double a, double lo, double hi) { return a < lo ? lo : (a > hi ? hi : a); }
template <int K> CHI_HD Dual<K> m_clip(const Dual<K>& a, double lo, double hi) {
  if (a.v < lo) return Dual<K>(lo);
  if (a.v > hi) return Dual<K>(hi);
  return a;
}

}  // namespace chi
