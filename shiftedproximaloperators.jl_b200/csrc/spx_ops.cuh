// spx_ops.cuh -- per-element functors of the separable shifted operators.
// Each `apply` restates one reference loop body, operation by operation
// (SURVEY.md Appendix A); the citation sits on each functor.
#pragma once
#include "spx_common.cuh"

namespace spx {

// ---- FP64 building blocks sized for the B200 FP64 pipe (64 lanes/clk/SM: at
// the HBM roofline a Float64 element has ~150 FP64 issue slots in total) -------

// Correctly rounded a / s for a warp-uniform divisor s, given y = RN(1/s)
// (Markstein: q faithful + exact FMA residual => the corrected quotient is the
// IEEE quotient; two correction steps, checked on the device against the IEEE
// division by spx_selftest_math / tests/test_gpu_math.py).  Falls back to the true division outside the
// range where every intermediate is a normal number.
__device__ __forceinline__ double div_uniform(double a, double s, double y) {
  const double q0 = a * y;
  const double r0 = __fma_rn(-s, q0, a);
  const double q1 = __fma_rn(r0, y, q0);
  const double r1 = __fma_rn(-s, q1, a);
  double q2 = __fma_rn(r1, y, q1);
  const double aa = fabs(a);
  if (!(aa > 1e-250 && aa < 1e250)) q2 = (aa == 0.0) ? q0 : __ddiv_rn(a, s);  // rare: a real branch
  return q2;
}

// sqrt(x) and 1/(2 sqrt(x)) for a normal positive x, branch-free: MUFU.RSQ64H seed
// (rsqrt.approx.ftz.f64, ~2^-22), two Goldschmidt rounds (g -> sqrt x, h -> 1/(2 sqrt x), each
// round squares the error), one final residual correction, which makes g the correctly
// rounded square root (same construction as the library's, minus its special-case branches).
__device__ __forceinline__ void sqrt_pair(double x, double& g, double& h) {
  double y;
  asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
  g = x * y;
  h = 0.5 * y;
  double r = __fma_rn(-g, h, 0.5);
  g = __fma_rn(g, r, g);
  h = __fma_rn(h, r, h);
  r = __fma_rn(-g, h, 0.5);
  g = __fma_rn(g, r, g);
  h = __fma_rn(h, r, h);
  const double d = __fma_rn(-g, g, x);
  g = __fma_rn(d, h, g);
}
// IEEE sqrt: straight-line for normal operands and for ±0 (a thresholded entry of the ψ(y) sum);
// subnormal / huge / negative / NaN take the library
static __device__ __noinline__ double sqrt_slow(double x) { return sqrt(x); }
__device__ __forceinline__ double sqrt_fast(double x) {
  const bool ok = x > 1e-290 && x < 1e290;
  if (!ok && x != 0.0) return sqrt_slow(x);
  double g, h;
  sqrt_pair(ok ? x : 1.0, g, h);
  return ok ? g : x;
}
__device__ __forceinline__ float sqrt_fast(float x) { return sqrtf(x); }

// RootNormLhalf closed form, operation by operation (the rare path of the kernels; the common one is the
// stationarity cubic below): `2*sign(z)/3*|z|*(1+cos(2π/3 - 2ϕ(z)/3))`, ϕ(z) = acos(νλ/4 (|z|/3)^(-3/2))
// (shiftedRootNormLhalf.jl:48,57).  The power, acos and cos run in Float64 for every R (the Float64 literals
// -3/2 and 2π/3 promote); the leading factors stay in R.
// t = c4 * w^(-3/2) to ~2 ulp (the reference's `^` is < 1 ulp, its product adds half an ulp): s = sqrt(w)
// correctly rounded, u = w*s, then the quotient c4/u by one Markstein step on the reciprocal estimate r^3.
static __device__ __noinline__ double lhalf_t_slow(double c4, double w) { return c4 * pow(w, -1.5); }
__device__ __forceinline__ double lhalf_t(double c4, double w) {
  if (!(w > 1e-200 && w < 1e200)) return lhalf_t_slow(c4, w);  // 0, Inf, NaN, denormal: library path
  double s, h;
  sqrt_pair(w, s, h);
  const double u = w * s;         // w^(3/2), <= 1 ulp
  const double r = h + h;         // ~ w^(-1/2)
  const double y = (r * r) * r;   // reciprocal estimate of u
  const double q0 = c4 * y;
  const double e2 = __fma_rn(-u, q0, c4);
  return __fma_rn(e2, y, q0);
}

// One Newton step on T3(x) = 4x^3 - 3x = rhs with a compensated residual: the
// square is split into p + e (exact), 4p - 3 is exact, so the residual carries
// one rounding of a quantity that is ~0 at the root; the slope is inverted in
// Float32 (error 1e-7: the step stays quadratic down to 1e-19).
__device__ __forceinline__ double cubic_newton(double x, double rhs) {
  const double p = x * x;
  const double e = __fma_rn(x, x, -p);
  const double u = __fma_rn(4.0, p, -3.0);
  double f = __fma_rn(x, u, -rhs);
  f = __fma_rn(4.0 * x, e, f);
  const float slope = (float)__fma_rn(12.0, p, -3.0);
  const double y = (double)__frcp_rn(slope);
  return __fma_rn(-f, y, x);
}

// G(t) = real(1 + cos(2π/3 - (2/3) acos(t + 0im)))   (shiftedRootNormLhalf.jl:57,
// shiftedRootNormLhalfBox.jl:92,106) without acos/cos:
//   t <= 1:  ψ = π/3 - acos(t)/3 ∈ [π/6, π/3] has cos 3ψ = -t, and 1 + cos 2ψ = 2 cos²ψ,
//            so G = 2 d² with d the root of 4d³ - 3d = -t in [1/2, √3/2];
//   t >  1:  acos(t + 0im) = -i acosh t, real(cos(a + ib)) = cos a cosh b, and
//            cosh(2x) = 2cosh²x - 1, so G = 1 + cos(2π/3) (2c² - 1) with c >= 1 the root of
//            4c³ - 3c = t.
// Start values come from Float32 (SFU) evaluations of the closed forms, then two
// compensated Newton steps in Float64 (a third next to the double root t -> 1).
// Measured against 200-bit mpmath: <= 2 ulp for t <= 0.9999, i.e. at the level of
// the reference's own acos/cos chain (tests/test_gpu_parity.py).
__device__ __forceinline__ double lhalf_G_real(double t) {  // 0 <= t <= 1
  // start value: d = cos(π/3 - acos(t)/3) is analytic in s = sqrt(1 - t) on [0, 1]; a degree-5
  // Float32 fit (|err| < 3e-7) is all two Newton steps need (error after two steps ~ K³ e⁴)
  float sf;
  asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(sf) : "f"((float)(1.0 - t)));
  float pf = 0.0011198767460882664f;
  pf = fmaf(pf, sf, -0.005838877987116575f);
  pf = fmaf(pf, sf, 0.01784452795982361f);
  pf = fmaf(pf, sf, -0.05533028766512871f);
  pf = fmaf(pf, sf, 0.40823012590408325f);
  pf = fmaf(pf, sf, 0.5000002384185791f);
  double d = (double)pf;
  d = cubic_newton(d, -t);
  d = cubic_newton(d, -t);
  if (t > 0.98) {
    d = cubic_newton(d, -t);
    d = cubic_newton(d, -t);
  }
  return 2.0 * (d * d);
}
// a / s for a warp-uniform divisor s: Float64 goes through div_uniform, Float32
// through the native (FP32-pipe) division
template <class R> struct UDiv;
template <> struct UDiv<double> {
  double s, y;
  bool fast;
  __host__ void set(double s_) {
    s = s_;
    y = 1.0 / s_;
    fast = std::isfinite(s_) && std::fabs(s_) > 1e-30 && std::fabs(s_) < 1e30;
  }
  __device__ __forceinline__ double operator()(double a) const { return fast ? div_uniform(a, s, y) : a / s; }
};
template <> struct UDiv<float> {
  float s;
  __host__ void set(float s_) { s = s_; }
  __device__ __forceinline__ float operator()(float a) const { return a / s; }
};

// ---- RootNormLhalf through its stationarity equation ------------------------------------------
// With t = (σλ/4)(|z|/3)^(-3/2) <= 1 the reference's value
//     (2/3)|z| (1 + cos(2π/3 - (2/3) acos t))                 (shiftedRootNormLhalf.jl:48,57)
// is s² where s is the largest root of  s³ - |z| s + σλ/2 = 0  (s = 2 sqrt(|z|/3) d turns this into
// 4d³ - 3d + t = 0, d = cos(π/3 - acos(t)/3)): it is the stationarity condition of
// y -> (y - |z|)²/(2σ) + λ sqrt(y) written in sqrt(y).  The kernels solve that cubic directly:
//   * start value in Float32 from the closed form (FP32 pipe + 3 SFU ops): rel. error ~1e-6;
//   * Newton in Float64 with fused residuals u = fma(s,s,-|z|), f = fma(s,u,a) (each a single
//     rounding of a quantity that vanishes at the root) and the Float32 reciprocal slope
//     (a stale slope only adds a term slope_err * step, far below 1 ulp after two steps).
// 7 FP64 operations instead of ~45 for pow + acos + cos.  Against 200-bit mpmath
// (tools/proto/check_lhalf_newton.py): <= 2 ulp for t <= 0.99 (<= 4 up to 0.998), where the
// reference's own pow/acos/cos chain is 3-40 ulp off; t in (0.998, 1.002) and operands outside the Float32-friendly
// range go through the operation-by-operation form (lhalf_mag_ref) in a non-inlined call.
struct LhalfStart {
  float t32;  // t in Float32 (rel. error ~5e-7)
  float s0;   // start value of s
  float inv;  // 1 / (3 s0² - |z|)
};
__device__ __forceinline__ float sqrt_approx(float x) {
  float r;
  asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}
__device__ __forceinline__ LhalfStart lhalf_start(float zf, float c4f) {
  LhalfStart st;
  const float w = zf * 0.33333334f;
  float r;
  asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(w));
  st.t32 = (c4f * r) * (r * r);
  // d = cos(π/3 - acos(t)/3) is analytic in sqrt(1 - t) on [0, 1]: degree-5 fit, |err| < 3e-7
  const float sf = sqrt_approx(fmaxf(1.0f - st.t32, 0.0f));
  float pf = 0.0011198767460882664f;
  pf = fmaf(pf, sf, -0.005838877987116575f);
  pf = fmaf(pf, sf, 0.01784452795982361f);
  pf = fmaf(pf, sf, -0.05533028766512871f);
  pf = fmaf(pf, sf, 0.40823012590408325f);
  pf = fmaf(pf, sf, 0.5000002384185791f);
  st.s0 = (2.0f * (w * r)) * pf;  // 2 sqrt(w) d
  const float slope = fmaf(3.0f * st.s0, st.s0, -zf);
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(st.inv) : "f"(slope));
  return st;
}
// s² for the largest root s; a = σλ/2.  NEAR1: t may approach 1 (Box form), where the slope
// az(4d² - 1) shrinks: one more step and a compensated last residual (s² = p + e, p - az by Fast2Sum) for
// 0.9 < t < 1.002 (beyond 1.002 the stationary point is not a candidate and the value is never used).
// SINGLE: the result is rounded to Float32 (R = Float32): one Float64 step already leaves an error of
// ~1e-11, far below a Float32 ulp (two where the slope shrinks).
template <bool NEAR1, bool SINGLE = false>
__device__ __forceinline__ double lhalf_newton(double az, double a, const LhalfStart& st) {
  const double inv = (double)st.inv;
  double s = (double)st.s0;
  double u = __fma_rn(s, s, -az);
  double f = __fma_rn(s, u, a);
  s = __fma_rn(-f, inv, s);
  if (SINGLE) {
    if (NEAR1 && st.t32 > 0.9f && st.t32 < 1.002f) {
      u = __fma_rn(s, s, -az);
      f = __fma_rn(s, u, a);
      s = __fma_rn(-f, inv, s);
      u = __fma_rn(s, s, -az);
      f = __fma_rn(s, u, a);
      s = __fma_rn(-f, inv, s);
    }
    return s * s;
  }
  if (NEAR1 && st.t32 > 0.9f && st.t32 < 1.002f) {
    u = __fma_rn(s, s, -az);
    f = __fma_rn(s, u, a);
    s = __fma_rn(-f, inv, s);
    const double p = s * s;
    const double e = __fma_rn(s, s, -p);
    u = p - az;
    const double ue = p - (u + az);
    f = __fma_rn(s, u, a);
    f = __fma_rn(s, e + ue, f);
    s = __fma_rn(-f, inv, s);
  } else {
    u = __fma_rn(s, s, -az);
    f = __fma_rn(s, u, a);
    s = __fma_rn(-f, inv, s);
  }
  return s * s;
}
// |z| and σλ/4 for which every Float32 intermediate above stays a normal number
__device__ __forceinline__ bool lhalf_f32_range(float zf) { return zf > 1e-9f && zf < 1e9f; }
inline bool lhalf_f32_range_host(double v) { return std::isfinite(v) && v > 1e-9 && v < 1e9; }

// operation-by-operation form of the magnitude (|z| > 0): ((2/3)|z|) G(t), t = c4 (|z|/3)^(-3/2)
template <class R> static __device__ __noinline__ double lhalf_mag_ref(R az, double c4) {
  const double t = lhalf_t(c4, (double)(az / R(3)));
  const R coef = (R(2) / R(3)) * az;
  return (double)coef * lhalf_G_real(t);
}

// ψ(y) accumulators (fused into the prox pass or run alone) --------------------
// kinds: SPX_H_L1 Σ|v|, SPX_H_L0 / SPX_H_INDBALLL0 Σ(v≠0), SPX_H_LHALF Σ√|v|
template <class R> __device__ __forceinline__ double h_term(int kind, R v) {
  if (kind == SPX_H_L1) return (double)jl_abs(v);
  if (kind == SPX_H_LHALF) return (double)sqrt_fast(jl_abs(v));  // IEEE sqrt in R (rootNormLhalf.jl:28)
  return (v != R(0)) ? 1.0 : 0.0;
}
// bad-flag encoding: max-reduced; "first failing index" is stored as 2^62 - i
__device__ __forceinline__ void flag_index(Partial& acc, long long i) {
  long long code = (1ll << 62) - i;
  acc.bad = code > acc.bad ? code : acc.bad;
}
__device__ __forceinline__ void flag_set(Partial& acc) { acc.bad = acc.bad > 1 ? acc.bad : 1; }

// ------------------------------------------------------------ ShiftedNormL1 --
// shiftedNormL1.jl:40-54
template <class R, bool PSI> struct ProxL1 {
  using Real = R;
  static constexpr int NIN = 3, UNROLL = 2;
  static constexpr bool OUT = true, ACC = PSI;
  const R* in[NIN];  // xk, sj, q
  R fill[NIN];
  R* y;
  R a;  // λσ
  __device__ __forceinline__ R apply(const R (&x)[NIN], long long, Partial& acc) const {
    R t = (-x[0]) - x[1];
    R o = jl_min(jl_max(t, x[2] - a), x[2] + a);
    if (PSI) acc.s += h_term(SPX_H_L1, (x[0] + x[1]) + o);
    return o;
  }
};
// shiftedNormL1.jl:60-75
template <class R, bool PSI> struct IproxL1 {
  using Real = R;
  static constexpr int NIN = 4, UNROLL = 2;
  static constexpr bool OUT = true, ACC = true;
  const R* in[NIN];  // xk, sj, g, d
  R fill[NIN];
  R* y;
  R lambda;
  __device__ __forceinline__ R apply(const R (&x)[NIN], long long i, Partial& acc) const {
    R t = (-x[0]) - x[1];
    R d = x[3];
    if (!(d > R(0))) flag_index(acc, i);  // @assert d[i] > 0  (:70)
    R c = (-x[2]) / d;
    R w = lambda / d;
    R o = jl_min(jl_max(t, c - w), c + w);
    if (PSI) acc.s += h_term(SPX_H_L1, (x[0] + x[1]) + o);
    return o;
  }
};

// ------------------------------------------------------------ ShiftedNormL0 --
// shiftedNormL0.jl:38-55
template <class R, bool PSI> struct ProxL0 {
  using Real = R;
  static constexpr int NIN = 3, UNROLL = 2;
  static constexpr bool OUT = true, ACC = PSI;
  const R* in[NIN];  // xk, sj, q
  R fill[NIN];
  R* y;
  R c;  // sqrt(2λσ)
  __device__ __forceinline__ R apply(const R (&x)[NIN], long long, Partial& acc) const {
    R xps = x[0] + x[1];
    R o = (jl_abs(xps + x[2]) <= c) ? -xps : x[2];
    if (PSI) acc.s += h_term(SPX_H_L0, xps + o);
    return o;
  }
};
// shiftedNormL0.jl:61-80
template <class R, bool PSI> struct IproxL0 {
  using Real = R;
  static constexpr int NIN = 4, UNROLL = 2;
  static constexpr bool OUT = true, ACC = true;
  const R* in[NIN];  // xk, sj, g, d
  R fill[NIN];
  R* y;
  R lambda;
  __device__ __forceinline__ R apply(const R (&x)[NIN], long long i, Partial& acc) const {
    R d = x[3];
    if (!(d > R(0))) flag_index(acc, i);  // @assert d[i] > 0  (:70)
    R ci = sqrt(R(2) * lambda * d);
    R xps = x[0] + x[1];
    R o = (jl_abs(d * xps - x[2]) <= ci) ? -xps : (-x[2]) / d;
    if (PSI) acc.s += h_term(SPX_H_L0, xps + o);
    return o;
  }
};

// shiftedRootNormLhalf.jl:41-63
template <class R, bool PSI> struct ProxLhalf {
  using Real = R;
#ifndef SPX_LH_UNROLL
#define SPX_LH_UNROLL 2
#endif
#ifndef SPX_LH_MINB
#define SPX_LH_MINB 4
#endif
  static constexpr int NIN = 3, UNROLL = SPX_LH_UNROLL, MINB = SPX_LH_MINB;
  static constexpr bool OUT = true, ACC = PSI;
  const R* in[NIN];  // xk, sj, q
  R fill[NIN];
  R* y;
  double p;    // 54^(1/3) (2νλ)^(2/3) / 4  (Float64)
  double c4;   // (double)(νλ/4), νλ/4 evaluated in R
  float c4f;   // (float)c4
  bool fast;   // c4 inside the Float32-friendly range (host-checked)
  __device__ __forceinline__ R apply(const R (&x)[NIN], long long, Partial& acc) const {
    const R xs = x[0] + x[1];
    const R z = x[2] + xs;  // ψ.sol[i]  (:50)
    const R az = jl_abs(z);
    // above the threshold t = νλ/4 (|z|/3)^(-3/2) lies in (0, 1/√2]: real branch only.  Every lane
    // runs the straight-line Newton (nearly every warp holds an element above the threshold).
    const bool above = !((double)az <= p);
    const float zf = (float)az;
    double mag = lhalf_newton<false, sizeof(R) == 4>((double)az, c4 + c4, lhalf_start(zf, c4f));
    if (above && !(fast && lhalf_f32_range(zf))) mag = lhalf_mag_ref<R>(az, c4);
    R o = above ? (R)copysign(mag, (double)z) : R(0);
    o = o - xs;
    if (PSI) acc.s += h_term(SPX_H_LHALF, xs + o);
    return o;
  }
};

// The three quotients the Box iprox! regimes with |d| >= eps need (shiftedNormL0Box.jl:179-188):
//   a = (-g)/d,  b = g/(d/2),  c = λ/(d/2).
// Reference form: three IEEE divisions.  Float64 fast form: ONE division (the correctly rounded
// reciprocal y = 1/d), a and λ/d by two Markstein corrections each (= the IEEE quotients, see
// div_uniform), and b = -2a, c = 2(λ/d): scaling by 2 commutes with rounding while nothing is subnormal
// and nothing overflows.  Elements outside that range take the reference form through one
// non-inlined call, so the straight-line path carries no fallback code.
template <class R> struct Quot3 { R a, b, c; };

template <class R> __device__ __forceinline__ Quot3<R> quot3_ref(R g, R d, R lambda) {
  Quot3<R> q;
  const R d_2 = d / R(2);
  q.a = (-g) / d;
  q.b = g / d_2;
  q.c = lambda / d_2;
  return q;
}
static __device__ __noinline__ Quot3<double> quot3_ref_call(double g, double d, double lambda) {
  return quot3_ref<double>(g, d, lambda);
}
__device__ __forceinline__ double markstein(double a, double d, double y) {
  const double q0 = a * y;
  const double r0 = __fma_rn(-d, q0, a);
  const double q1 = __fma_rn(r0, y, q0);
  const double r1 = __fma_rn(-d, q1, a);
  const double q2 = __fma_rn(r1, y, q1);
  return a == 0.0 ? q0 : q2;  // ±0 numerators: the product already carries the IEEE sign
}
__device__ __forceinline__ Quot3<float> quot3(float g, float d, float lambda, bool) {
  return quot3_ref<float>(g, d, lambda);  // FP32 divisions run on the FP32 pipe
}
__device__ __forceinline__ Quot3<double> quot3(double g, double d, double lambda, bool lam_ok) {
  const double ad = fabs(d), ag = fabs(g);
  const bool fast = lam_ok && ad > 1e-100 && ad < 1e100 && (g == 0.0 || (ag > 1e-150 && ag < 1e150));
  if (!fast) return quot3_ref_call(g, d, lambda);
  const double y = 1.0 / d;
  Quot3<double> q;
  q.a = markstein(-g, d, y);
  q.b = -2.0 * q.a;
  q.c = 2.0 * markstein(lambda, d, y);
  return q;
}

// iprox_zero (ShiftedProximalOperators.jl:217-236) reusing an already computed (-g)/d
template <class R> __device__ __forceinline__ R iprox_zero_q(R d, R g, R l, R u, R neg_g_over_d) {
  const R eps = Eps<R>::value;
  const R r1 = jl_min(jl_max(neg_g_over_d, l), u);
  const R d_2 = d * R(0.5);
  const R val_l = d_2 * (l * l) + g * l;
  const R val_u = d_2 * (u * u) + g * u;
  const R r2 = (val_l < val_u) ? l : u;
  const R r3 = (g > R(0)) ? l : ((g < R(0)) ? u : R(0));
  return (d > eps) ? r1 : ((d < -eps) ? r2 : r3);
}

// ------------------------------------------------------------------ Box types --
// inputs: 0 xk, 1 sj, 2 q (or g), [3 d], then l, u (nullable -> scalar fill)
template <class R> struct BoxPsi {
  int kind;  // SPX_H_L1 / _L0 / _LHALF
  // Box ψ(y): h over the selected entries + Inf if any sj+y ∉ [l-ϵ, u+ϵ]
  // (shiftedNormL1Box.jl:70-82)
  __device__ __forceinline__ void add(Partial& acc, bool selected, R xk, R sj, R y, R l, R u) const {
    if (selected) acc.s += h_term(kind, (xk + sj) + y);
    const R e = Eps<R>::sqrt_value;
    R w = sj + y;
    if (!((l - e <= w) && (w <= u + e))) flag_set(acc);
  }
};

// shiftedNormL1Box.jl:89-125
// resident CTAs per SM asked of ptxas: the fused-ψ forms spill under the 64 registers of four CTAs (measured at
// n = 2^28 Float64: L0Box+ψ 84.3 % -> 90.6 %, L1Box+ψ 88.2 % -> 92.1 %, L0Box 88.2 % -> 93.2 %, L1Box 92.5 % -> 94.8 % of
// HBM peak with three; LhalfBox keeps four: 85.6 % vs 76.8 %)
#ifndef SPX_PSI_MINB
#define SPX_PSI_MINB 3
#endif
#ifndef SPX_BOX_MINB
#define SPX_BOX_MINB 3
#endif
template <class R, bool PSI> struct ProxL1Box {
  using Real = R;
  static constexpr int NIN = 5, UNROLL = 2, MINB = PSI ? SPX_PSI_MINB : SPX_BOX_MINB;
  static constexpr bool OUT = true, ACC = PSI;
  static constexpr bool SPLIT_NULL = true;  // l, u: both vectors or both scalars at compile time
  const R* in[NIN];  // xk, sj, q, l, u
  R fill[NIN];
  R* y;
  DevSel sel;
  R sl;  // σλ
  __device__ __forceinline__ R apply(const R (&x)[NIN], long long i, Partial& acc) const {
    const R xi = x[0], si = x[1], qi = x[2], li = x[3], ui = x[4];
    const bool s = sel.has(i);
    R o;
    if (s) {
      R xs = xi + si;
      R xsq = xs + qi;
      R yi = (xsq <= -sl) ? qi + sl : ((xsq >= sl) ? qi - sl : -xs);
      o = jl_min(jl_max(yi, li - si), ui - si);
    } else {
      o = prox_zero(qi, li - si, ui - si);
    }
    if (PSI) BoxPsi<R>{SPX_H_L1}.add(acc, s, xi, si, o, li, ui);
    return o;
  }
};

// The four quotients of the L1Box iprox! regimes with |d| > eps (shiftedNormL1Box.jl:161-171):
//   g/(d/2) = 2 (g/d),  λ/(d/2) = 2 (λ/d),  a1 = (-(g+λ))/d,  a2 = (λ-g)/d
// from one reciprocal (see Quot3).  Sums and differences of g and λ are 0 or at least 2^-53 of the larger,
// so the range test on g and λ covers all four numerators.
template <class R> struct Quot4 { R g2, l2, a1, a2; };
template <class R> __device__ __forceinline__ Quot4<R> quot4_ref(R g, R d, R lambda) {
  Quot4<R> q;
  const R d_2 = d / R(2);
  q.g2 = g / d_2;
  q.l2 = lambda / d_2;
  q.a1 = (-(g + lambda)) / d;
  q.a2 = (lambda - g) / d;
  return q;
}
static __device__ __noinline__ Quot4<double> quot4_ref_call(double g, double d, double lambda) {
  return quot4_ref<double>(g, d, lambda);
}
__device__ __forceinline__ Quot4<float> quot4(float g, float d, float lambda, bool) {
  return quot4_ref<float>(g, d, lambda);
}
__device__ __forceinline__ Quot4<double> quot4(double g, double d, double lambda, bool lam_ok) {
  const double ad = fabs(d), ag = fabs(g);
  const bool fast = lam_ok && ad > 1e-100 && ad < 1e100 && (g == 0.0 || (ag > 1e-100 && ag < 1e100));
  if (!fast) return quot4_ref_call(g, d, lambda);
  const double y = 1.0 / d;
  Quot4<double> q;
  q.g2 = 2.0 * markstein(g, d, y);
  q.l2 = 2.0 * markstein(lambda, d, y);
  q.a1 = markstein(-(g + lambda), d, y);
  q.a2 = markstein(lambda - g, d, y);
  return q;
}

// shiftedNormL1Box.jl:131-225.  Straight-line like IproxL0Box below: the three regimes are evaluated
// for every element and selected.  `x < min(a, b)` is written `x < a && x < b` (same truth table,
// NaNs included, since Base.min propagates NaN and every comparison with NaN is false).
#ifndef SPX_IPB_MINB
#define SPX_IPB_MINB 3
#endif
template <class R, bool PSI> struct IproxL1Box {
  using Real = R;
  static constexpr int NIN = 6, UNROLL = 1, MINB = SPX_IPB_MINB;
  static constexpr bool OUT = true, ACC = PSI;
  static constexpr bool SPLIT_NULL = true;  // l, u: both vectors or both scalars at compile time
  const R* in[NIN];  // xk, sj, g, d, l, u
  R fill[NIN];
  R* y;
  DevSel sel;
  R lambda;
  bool lam_ok;  // λ is 0 or a normal number far from the range limits (host-checked)
  __device__ __forceinline__ R apply(const R (&x)[NIN], long long i, Partial& acc) const {
    const R xi = x[0], si = x[1], gi = x[2], di = x[3], li = x[4], ui = x[5];
    const R eps = Eps<R>::value;
    const bool s = (sel.kind == SPX_SEL_ALL) || sel.has(i);
    const R xs = xi + si, mxs = -xs;
    const R left = li - si, right = ui - si;
    // ---- |d| <= eps  (:152-159)
    const R yA = (jl_abs(gi) <= lambda) ? jl_min(jl_max(left, mxs), right) : ((gi > R(0)) ? left : right);
    // ---- |d| > eps: one reciprocal serves the four quotients
    const bool small_d = jl_abs(di) <= eps;
    const Quot4<R> qt = quot4(gi, small_d ? R(1) : di, lambda, lam_ok);
    const R lx = li + xi, ux = ui + xi;
    const R fi2_di = qt.g2 - R(2) * xs;
    const R l2_di = qt.l2;
    const R vl = lx * lx + fi2_di * lx + l2_di * jl_abs(lx);
    const R vr = ux * ux + fi2_di * ux + l2_di * jl_abs(ux);
    R yB;
    {  // d > eps  (:161-198)
      const R edge = (vl < vr) ? left : right;
      const R a1 = qt.a1, a2 = qt.a2;
      const bool in1 = (left <= a1) && (a1 <= right);
      const bool in2 = (left <= a2) && (a2 <= right);
      const R v1 = xs + a1, v2 = xs + a2;
      const R val1 = v1 * v1 + fi2_di * v1 + l2_di * jl_abs(v1);
      const R val2 = v2 * v2 + fi2_di * v2 + l2_di * jl_abs(v2);
      const bool c1 = in1 && (val1 < vl) && (val1 < vr);
      const bool c2 = in2 && (val2 < vl) && (val2 < vr) && (!in1 || (val2 < val1));
      const bool c0 = (R(0) < vl) && (R(0) < vr) && (!in1 || (R(0) < val1)) && (!in2 || (R(0) < val2));
      R mid = c1 ? a1 : edge;
      mid = c2 ? a2 : mid;
      mid = c0 ? mxs : mid;
      const R pos = in1 ? a1 : edge;  // lx >= 0
      const R neg = in2 ? a2 : edge;  // 0 >= ux
      yB = (lx >= R(0)) ? pos : ((R(0) >= ux) ? neg : mid);
    }
    R yC;
    {  // d < -eps  (:200-218)
      yC = (vl > vr) ? left : right;
      if ((li <= -xi) && (-xi <= ui) && (R(0) > vl) && (R(0) > vr)) yC = mxs;
    }
    R yi = small_d ? yA : ((di > eps) ? yB : yC);
    if (!s) yi = iprox_zero(di, gi, left, right);
    if (PSI) BoxPsi<R>{SPX_H_L1}.add(acc, s, xi, si, yi, li, ui);
    return yi;
  }
};

// shiftedNormL0Box.jl:89-131
template <class R, bool PSI> struct ProxL0Box {
  using Real = R;
  static constexpr int NIN = 5, UNROLL = 2, MINB = PSI ? SPX_PSI_MINB : SPX_BOX_MINB;
  static constexpr bool OUT = true, ACC = PSI;
  static constexpr bool SPLIT_NULL = true;  // l, u: both vectors or both scalars at compile time
  const R* in[NIN];  // xk, sj, q, l, u
  R fill[NIN];
  R* y;
  DevSel sel;
  R c;  // 2λσ
  __device__ __forceinline__ R apply(const R (&x)[NIN], long long i, Partial& acc) const {
    const R xi = x[0], si = x[1], qi = x[2], li = x[3], ui = x[4];
    const bool s = sel.has(i);
    const R sq = si + qi;
    R o;
    if (s) {
      const R xs = xi + si;
      const R xsq = xs + qi;
      const R dl = li - sq, du = ui - sq;
      const R val_left = dl * dl + ((xi == -li) ? R(0) : c);
      const R val_right = du * du + ((xi == -ui) ? R(0) : c);
      o = (val_left < val_right) ? (li - si) : (ui - si);
      R val_min = jl_min(val_left, val_right);
      if ((li <= -xi) && (-xi <= ui)) {
        R val_0 = xsq * xsq;
        if (val_0 < val_min) o = -xs;
        val_min = jl_min(val_0, val_min);
      }
      if ((li <= sq) && (sq <= ui)) {
        R val_xsq = (xsq == R(0)) ? R(0) : c;
        if (val_xsq < val_min) o = qi;
      }
    } else {
      o = prox_zero(qi, li - si, ui - si);
    }
    if (PSI) BoxPsi<R>{SPX_H_L0}.add(acc, s, xi, si, o, li, ui);
    return o;
  }
};

// shiftedNormL0Box.jl:137-231.  The three regimes (|d| < eps, d >= eps, d <= -eps) are evaluated
// straight-line and selected: with 64 elements per warp and mixed-sign d every warp needs all of them,
// so predication beats divergence, and the two lanes of a 128-bit pair interleave.
template <class R, bool PSI> struct IproxL0Box {
  using Real = R;
  static constexpr int NIN = 6, UNROLL = 1, MINB = SPX_IPB_MINB;
  static constexpr bool OUT = true, ACC = PSI;
  static constexpr bool SPLIT_NULL = true;  // l, u: both vectors or both scalars at compile time
  const R* in[NIN];  // xk, sj, g, d, l, u
  R fill[NIN];
  R* y;
  DevSel sel;
  R lambda;
  bool lam_ok;  // λ is 0 or a normal number far from the range limits (host-checked)
  __device__ __forceinline__ R apply(const R (&x)[NIN], long long i, Partial& acc) const {
    const R xi = x[0], si = x[1], gi = x[2], di = x[3], li = x[4], ui = x[5];
    const R eps = Eps<R>::value;
    const bool s = (sel.kind == SPX_SEL_ALL) || sel.has(i);
    const R xs = xi + si, mxs = -xs;
    const R left = li - si, right = ui - si;
    const bool zero_in = (li <= -xi) && (-xi <= ui);
    const R lx = li + xi, ux = ui + xi;
    // ---- regime |d| < eps  (:155-177) ----
    R y1;
    {
      const R vl = gi * left + ((xi == -li) ? R(0) : lambda);
      const R vr = gi * right + ((xi == -ui) ? R(0) : lambda);
      const bool gpos = gi > R(0), gneg = gi < R(0);
      // a NaN g takes no branch in the reference (y[i] untouched); NaN is written here
      const R nan = gi - gi;  // NaN exactly when g is NaN, the only case in which it is used
      const R val_min = gpos ? vl : (gneg ? vr : nan);
      R yy = gpos ? left : (gneg ? right : nan);
      const R val_0 = (-gi) * xs;
      if (zero_in && (val_0 < val_min)) yy = mxs;
      y1 = (gi == R(0)) ? (zero_in ? mxs : R(0)) : yy;
    }
    // ---- regimes |d| >= eps  (:179-224): one reciprocal serves every quotient ----
    const bool small_d = jl_abs(di) < eps;
    const Quot3<R> qt = quot3(gi, small_d ? R(1) : di, lambda, lam_ok);
    const R aq_y = qt.a;    // (-g)/d
    const R gi2_di = qt.b;  // g/(d/2)
    const R l2_di = qt.c;   // λ/(d/2)
    const R fi2_di = gi2_di - R(2) * xs;
    const R val_left = (lx == R(0)) ? R(0) : (lx * lx + fi2_di * lx + l2_di);
    const R val_right = (ux == R(0)) ? R(0) : (ux * ux + fi2_di * ux + l2_di);
    R y2;
    {  // d >= eps  (:189-209)
      const R aq_v = aq_y + xs;
      const bool inside = (lx <= aq_v) && (aq_v <= ux);
      const R sq = aq_v * aq_v;
      const R vin = (aq_v == R(0)) ? -sq : (-sq + l2_di);
      // 0 < min(val_left, val_right)  <=>  0 < val_left && 0 < val_right  (Base.min propagates NaN)
      const bool pos_min = inside ? (R(0) < vin) : ((R(0) < val_left) && (R(0) < val_right));
      y2 = inside ? aq_y : ((val_left < val_right) ? left : right);
      if (zero_in && pos_min) y2 = mxs;
    }
    R y3;
    {  // d <= -eps  (:211-223)
      y3 = (val_left > val_right) ? left : right;
      if (zero_in && (R(0) > val_left) && (R(0) > val_right)) y3 = mxs;
    }
    R yi = small_d ? y1 : ((di >= eps) ? y2 : y3);
    if (!s) yi = iprox_zero_q(di, gi, left, right, aq_y);
    if (PSI) BoxPsi<R>{SPX_H_L0}.add(acc, s, xi, si, yi, li, ui);
    return yi;
  }
};

// shiftedRootNormLhalfBox.jl:86-120 -- the operation-by-operation form of one selected element: four
// candidate objectives RNorm(tt) = (tt - q)²/2/σ + λ sqrt|tt + xs| with IEEE divisions and square roots,
// then `findmin` (first minimal index, NaN counts as minimal).  The streaming kernel calls it only for
// the elements its Float32 filter cannot decide (below).
template <class R> struct LhalfBoxExact {
  R lambda;
  double c4;  // (double)(σλ/4), σλ/4 evaluated in R
  UDiv<R> by3, by_sigma;
  UDiv<double> by_sigma64;
};
template <class R>
static __device__ __noinline__ R lhalfbox_exact(R xi, R si, R qi, R li, R ui, const LhalfBoxExact<R> k) {
  const R xs = xi + si;  // ψ.sol[i]  (:94)
  const R xsq = xs + qi;
  const R axsq = jl_abs(xsq);
  const double t = lhalf_t(k.c4, (double)k.by3(axsq));
  const R coef = (jl_sign(xsq) * (R(2) / R(3))) * axsq;  // == ((2 sign) / 3) |xsq| bit for bit
  // Candidate 4 (`val - xs`) only matters on the real branch t <= 1.  For t > 1 (|xsq| below
  // 3 (σλ/4)^(2/3)) the objective RNorm has no stationary point besides the kink at tt = -xs: it
  // decreases towards the kink from both sides, so whatever real(val) is, an endpoint or the kink is
  // never worse and `findmin` (strict improvement, candidate 4 last) cannot pick it; t = NaN/Inf
  // make val NaN/Inf, which the reference's range test rejects as well.  The complex-branch value is
  // therefore never evaluated here ("LhalfBox candidate 4").
  const bool real_branch = t <= 1.0;
  const double val = (double)coef * lhalf_G_real(real_branch ? t : 1.0);
  // x/2 == x*0.5 exactly
  const R left = li - si, right = ui - si, mxs = -xs;
  const R dl = left - qi, dr = right - qi;
  const double inf = __longlong_as_double(0x7ff0000000000000ll);
  const double c0 = (double)(k.by_sigma((dl * dl) * R(0.5)) + k.lambda * sqrt_fast(jl_abs(left + xs)));
  const double c1 = (double)(k.by_sigma((dr * dr) * R(0.5)) + k.lambda * sqrt_fast(jl_abs(right + xs)));
  // tt = -xs: tt - q = -(xs + q), and tt + xs is 0 (NaN for a non-finite xs) where sqrt|.| == |.|
  const double c2v = (double)(k.by_sigma((xsq * xsq) * R(0.5)) + k.lambda * jl_abs(mxs + xs));
  const double c2 = ((li <= -xi) && (-xi <= ui)) ? c2v : inf;
  const double vmx = val - (double)xi;
  const double cand4 = val - (double)xs;  // Float64 (val is Float64)
  const double dq4 = cand4 - (double)qi;
  const double c3v = k.by_sigma64((dq4 * dq4) * 0.5) + (double)k.lambda * sqrt_fast(fabs(cand4 + (double)xs));
  const double c3 = (real_branch && ((double)li <= vmx) && (vmx <= (double)ui)) ? c3v : inf;
  // findmin over (c0, c1, c2, c3): the values are NaN, +Inf or >= +0, so Base.isless reduces to `<`
  // plus the NaN rule.
  int a = 0;
  double fm = c0;
  if ((c1 < fm) || ((c1 != c1) && (fm == fm))) { fm = c1; a = 1; }
  if ((c2 < fm) || ((c2 != c2) && (fm == fm))) { fm = c2; a = 2; }
  if ((c3 < fm) || ((c3 != c3) && (fm == fm))) { fm = c3; a = 3; }
  return a == 0 ? left : (a == 1 ? right : (a == 2 ? mxs : (R)cand4));
}

// The streaming form.  `findmin` only needs the ORDER of the four objectives, so they are first
// evaluated in Float32 (FP32 pipe + SFU square roots; the differences tt - q and tt + xs are formed
// in R and rounded once, so every objective carries a relative error < 4e-7).  When the smallest and
// the second smallest differ by more than 4e-6 (ten times that bound) the exact order is decided;
// otherwise -- exact ties, NaNs, values outside the Float32 range, t within 0.2% of the real/complex
// boundary -- the element goes through lhalfbox_exact.  The selected point itself (an edge, the
// kink, or the stationary point from the Float64 Newton) never passes through Float32.
template <class R, bool PSI> struct ProxLhalfBox {
  using Real = R;
#ifndef SPX_LHB_UNROLL
#define SPX_LHB_UNROLL 1
#endif
#ifndef SPX_LHB_MINB
#define SPX_LHB_MINB 4
#endif
#ifndef SPX_LHB_STAGES
#define SPX_LHB_STAGES 0
#endif
#ifndef SPX_LHB_PSI_MINB
#define SPX_LHB_PSI_MINB 4
#endif
  static constexpr int NIN = 5, UNROLL = SPX_LHB_UNROLL, MINB = PSI ? SPX_LHB_PSI_MINB : SPX_LHB_MINB, STAGES = SPX_LHB_STAGES;
  static constexpr bool OUT = true, ACC = PSI;
  static constexpr bool SPLIT_NULL = true;  // l, u: both vectors or both scalars at compile time
  const R* in[NIN];  // xk, sj, q, l, u
  R fill[NIN];
  R* y;
  DevSel sel;
  LhalfBoxExact<R> k;
  float kf;    // 1/(2σ)
  float lamf;  // λ
  float c4f;   // σλ/4
  double a2;   // σλ/2 (= 2 c4), the constant term of the cubic
  float a2f;   // the same in Float32 (exact when R = Float32: c4 is σλ/4 evaluated in R)
  __device__ __forceinline__ const LhalfBoxExact<float>& kexact_f32() const {
    return reinterpret_cast<const LhalfBoxExact<float>&>(k);  // only called when R = Float32
  }
  bool fast;   // σ, λ, σλ/4 inside the Float32-friendly range (host-checked)
  // R = Float32.  The reference evaluates candidate 4 in Float64 (`val - xs`, the range test `l <= val - xk <= u`) and
  // rounds once when it stores into y.  Here val is carried as an unevaluated sum of two Float32 numbers (error
  // ~1e-11 |val|, that of the one-step Float64 form this replaces): s0² = p + e exactly (one FMA), p - |z| is exact
  // (Sterbenz: s² ∈ [0.52 |z|, |z|] for t <= 0.9), so the residual of the cubic loses nothing to cancellation; one
  // Newton step with the Float32 reciprocal slope, s² = p + (e + 2 s0 δ + δ²).  `val - xs` is a TwoSum.
  // No Float64 instruction and no Float32 <-> Float64 conversion on this path (they share the 16-lane XU pipe with the
  // square roots: 12 conversions + 6 SFU operations per element made the Float32 kernel slower per byte than the
  // Float64 one).  t32 in (0.9, 1.002) -- slope of the cubic shrinking, 0.5 % of the bench's elements -- takes the
  // three-step Float64 Newton; a range test within 4 ulps of a bound goes to lhalfbox_exact.
  __device__ __forceinline__ float apply_f32(const float (&x)[NIN], long long i, Partial& acc) const {
    const float xi = x[0], si = x[1], qi = x[2], li = x[3], ui = x[4];
    const bool s = (sel.kind == SPX_SEL_ALL) || sel.has(i);
    const float xs = xi + si;  // ψ.sol[i]  (:94)
    const float xsq = xs + qi;
    const float zf = fabsf(xsq);
    const float left = li - si, right = ui - si, mxs = -xs;
    const LhalfStart st = lhalf_start(zf, c4f);
    const bool real_branch = st.t32 <= 0.998f;
    bool hard = !(fast && lhalf_f32_range(zf)) || !(st.t32 <= 0.998f || st.t32 >= 1.002f);
    float mh, ml;  // |val| = mh + ml
    if (st.t32 > 0.9f && st.t32 < 1.002f) {
      const double mag = lhalf_newton<true, true>((double)zf, a2, st);
      mh = (float)mag;
      ml = (float)(mag - (double)mh);
    } else {
      const float s0 = st.s0;
      const float p = s0 * s0;
      const float e = fmaf(s0, s0, -p);
      const float uh = p - zf;
      const float t1 = s0 * uh;
      const float t1e = fmaf(s0, uh, -t1);
      const float f = (t1 + a2f) + fmaf(s0, e, t1e);
      const float d = -f * st.inv;
      const float m1 = fmaf(d, d, fmaf(s0 + s0, d, e));
      mh = p + m1;
      ml = (p - mh) + m1;  // Fast2Sum: |p| >= |m1|
    }
    // val = sign(xsq) (mh + ml); cand4 = val - xs and vmx = val - xk as two-term sums
    const float vh = copysignf(mh, xsq), vl = (xsq < 0.f) ? -ml : ml;
    const float th = vh - xs, tb = th - vh;
    const float cl = ((vh - (th - tb)) + (-xs - tb)) + vl;
    const float cand4 = th + cl;  // `val - xs` rounded to R once
    // li <= val - xk <= ui in Float64: val - xk = (val - xs) + sj; decided by the Float32 value of that (off by at
    // most half an ulp of each of the two terms) unless it sits within 4 ulps of a bound
    const float vmx = cand4 + si;
    const float guard = 2.4e-7f * (fabsf(vmx) + fabsf(cand4));
    hard = hard || (fabsf(vmx - li) <= guard) || (fabsf(vmx - ui) <= guard);
    const bool in4 = real_branch && (li <= vmx) && (vmx <= ui);
    const bool zero_in = (li <= -xi) && (-xi <= ui);
    // Float32 objectives (as in the generic form below)
    const float inff = __int_as_float(0x7f800000);
    const float dlf = left - qi, drf = right - qi;
    const float alf = fabsf(left + xs), arf = fabsf(right + xs);
    const float dq4f = (th - qi) + cl;
    const float c0 = fmaf(dlf * dlf, kf, lamf * sqrt_approx(alf));
    const float c1 = fmaf(drf * drf, kf, lamf * sqrt_approx(arf));
    const float c2 = zero_in ? (zf * zf) * kf : inff;
    const float c3 = in4 ? fmaf(dq4f * dq4f, kf, lamf * sqrt_approx(mh)) : inff;
    int a = 0;
    float m = c0;
    if (c1 < m) { m = c1; a = 1; }
    if (c2 < m) { m = c2; a = 2; }
    if (c3 < m) { m = c3; a = 3; }
    const float lo01 = fminf(c0, c1), hi01 = fmaxf(c0, c1);
    const float lo23 = fminf(c2, c3), hi23 = fmaxf(c2, c3);
    const float m2 = fminf(fmaxf(lo01, lo23), fminf(hi01, hi23));  // second smallest
    const bool clear = (m2 > m * 1.000004f) && (m > 1e-30f);
    // the objectives are >= +0, +Inf or NaN: their sum is NaN exactly when one of them is
    const float csum = (c0 + c1) + (c2 + c3);
    hard = hard || !clear || (csum != csum);
    float o = left;
    o = (a == 1) ? right : o;
    o = (a == 2) ? mxs : o;
    o = (a == 3) ? cand4 : o;
    if (hard) o = lhalfbox_exact<float>(xi, si, qi, li, ui, kexact_f32());
    if (!s) o = prox_zero(qi, left, right);
    if (PSI) BoxPsi<float>{SPX_H_LHALF}.add(acc, s, xi, si, o, li, ui);
    return o;
  }
  __device__ __forceinline__ R apply(const R (&x)[NIN], long long i, Partial& acc) const {
    if constexpr (sizeof(R) == 4) return apply_f32(x, i, acc);
    const R xi = x[0], si = x[1], qi = x[2], li = x[3], ui = x[4];
    const bool s = (sel.kind == SPX_SEL_ALL) || sel.has(i);
    const R xs = xi + si;  // ψ.sol[i]  (:94)
    const R xsq = xs + qi;
    const R axsq = jl_abs(xsq);
    const R left = li - si, right = ui - si, mxs = -xs;
    // stationary point (candidate 4)
    const float zf = (float)axsq;
    const LhalfStart st = lhalf_start(zf, c4f);
    const double mag = lhalf_newton<true, sizeof(R) == 4>((double)axsq, a2, st);
    const double val = copysign(mag, (double)xsq);
    const bool real_branch = st.t32 <= 0.998f;
    bool hard = !(fast && lhalf_f32_range(zf)) || !(st.t32 <= 0.998f || st.t32 >= 1.002f);
    const double vmx = val - (double)xi;
    const double cand4 = val - (double)xs;
    const double dq4 = cand4 - (double)qi;
    const double ar4 = cand4 + (double)xs;
    const bool in4 = real_branch && ((double)li <= vmx) && (vmx <= (double)ui);
    const bool zero_in = (li <= -xi) && (-xi <= ui);
    // Float32 objectives
    const float inff = __int_as_float(0x7f800000);
    const float dlf = (float)(left - qi), drf = (float)(right - qi);
    const float alf = fabsf((float)(left + xs)), arf = fabsf((float)(right + xs));
    const float dq4f = (float)dq4, a4f = fabsf((float)ar4);
    const float c0 = fmaf(dlf * dlf, kf, lamf * sqrt_approx(alf));
    const float c1 = fmaf(drf * drf, kf, lamf * sqrt_approx(arf));
    const float c2 = zero_in ? (zf * zf) * kf : inff;
    const float c3 = in4 ? fmaf(dq4f * dq4f, kf, lamf * sqrt_approx(a4f)) : inff;
    int a = 0;
    float m = c0;
    if (c1 < m) { m = c1; a = 1; }
    if (c2 < m) { m = c2; a = 2; }
    if (c3 < m) { m = c3; a = 3; }
    const float lo01 = fminf(c0, c1), hi01 = fmaxf(c0, c1);
    const float lo23 = fminf(c2, c3), hi23 = fmaxf(c2, c3);
    const float m2 = fminf(fmaxf(lo01, lo23), fminf(hi01, hi23));  // second smallest
    const bool clear = (m2 > m * 1.000004f) && (m > 1e-30f);
    const bool nonan = (c0 == c0) && (c1 == c1) && (c2 == c2) && (c3 == c3);
    hard = hard || !clear || !nonan;
    R o = left;
    o = (a == 1) ? right : o;
    o = (a == 2) ? mxs : o;
    o = (a == 3) ? (R)cand4 : o;
    if (hard) o = lhalfbox_exact<R>(xi, si, qi, li, ui, k);
    if (!s) o = prox_zero(qi, left, right);
    if (PSI) BoxPsi<R>{SPX_H_LHALF}.add(acc, s, xi, si, o, li, ui);
    return o;
  }
};

// ------------------------------------------------------------------- values --
// generic ψ(y) = h((xk + sj) + y)  ShiftedProximalOperators.jl:51-54
template <class R> struct ValueSep {
  using Real = R;
  static constexpr int NIN = 3, UNROLL = 4;
  static constexpr bool OUT = false, ACC = true;
  const R* in[NIN];  // xk, sj, y
  R fill[NIN];
  R* y;
  int kind;
  __device__ __forceinline__ R apply(const R (&x)[NIN], long long, Partial& acc) const {
    acc.s += h_term(kind, (x[0] + x[1]) + x[2]);
    return R(0);
  }
};
// Box ψ(y) (membership weights; a list with duplicates goes through ValueGather)
template <class R> struct ValueBox {
  using Real = R;
  static constexpr int NIN = 5, UNROLL = 2, MINB = 4;
  static constexpr bool OUT = false, ACC = true;
  static constexpr bool SPLIT_NULL = true;  // l, u: both vectors or both scalars at compile time
  const R* in[NIN];  // xk, sj, y, l, u
  R fill[NIN];
  R* y;
  DevSel sel;
  int kind;
  bool weigh;  // false -> feasibility sweep only
  __device__ __forceinline__ R apply(const R (&x)[NIN], long long i, Partial& acc) const {
    BoxPsi<R>{kind}.add(acc, weigh && sel.has(i), x[0], x[1], x[2], x[3], x[4]);
    return R(0);
  }
};
// ShiftedNormL1B2 ψ(y): Σ|xk+sj+y| and Σ(sj+y)²  (shiftedNormL1B2.jl:32)
template <class R> struct ValueL1B2 {
  using Real = R;
  static constexpr int NIN = 3, UNROLL = 4;
  static constexpr bool OUT = false, ACC = true;
  const R* in[NIN];  // xk, sj, y
  R fill[NIN];
  R* y;
  __device__ __forceinline__ R apply(const R (&x)[NIN], long long, Partial& acc) const {
    acc.s += (double)jl_abs((x[0] + x[1]) + x[2]);
    double w = (double)(x[1] + x[2]);
    acc.s2 += w * w;
    return R(0);
  }
};
// BInf ψ(y): w = sj + y; IndBallLinf(1.1Δ)(w) (strict, Float64 radius); v = w + xk
// (shiftedIndBallL0BInf.jl:44-49)
template <class R> struct ValueBinfCount {
  using Real = R;
  static constexpr int NIN = 3, UNROLL = 4;
  static constexpr bool OUT = false, ACC = true;
  const R* in[NIN];  // xk, sj, y
  R fill[NIN];
  R* y;
  double rad;  // 1.1 * Δ in Float64
  __device__ __forceinline__ R apply(const R (&x)[NIN], long long, Partial& acc) const {
    R w = x[1] + x[2];
    if ((double)w < -rad || (double)w > rad) flag_set(acc);
    acc.s += ((w + x[0]) != R(0)) ? 1.0 : 0.0;
    return R(0);
  }
};

}  // namespace spx
