// spx_ops.cuh -- per-element functors of the separable shifted operators.
// Each `apply` restates one reference loop body, operation by operation
// (SURVEY.md Appendix A); the citation sits on each functor.
#pragma once
#include "spx_common.cuh"

namespace spx {

// ψ(y) accumulators (fused into the prox pass or run alone) --------------------
// kinds: SPX_H_L1 Σ|v|, SPX_H_L0 / SPX_H_INDBALLL0 Σ(v≠0), SPX_H_LHALF Σ√|v|
template <class R> __device__ __forceinline__ double h_term(int kind, R v) {
  if (kind == SPX_H_L1) return (double)jl_abs(v);
  if (kind == SPX_H_LHALF) return (double)sqrt(jl_abs(v));  // sqrt in R (rootNormLhalf.jl:28)
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

// ---------------------------------------------------- RootNormLhalf closed form
// `2*sign(z)/3*|z|*(1+cos(2π/3 - 2ϕ(z)/3))`, ϕ(z) = acos(νλ/4 (|z|/3)^(-3/2))
// shiftedRootNormLhalf.jl:48,57.  The power, acos and cos run in Float64 for
// every R (Float64 literals -3/2, 2π/3 promote); the leading factors stay in R.
constexpr double kTwoPiOver3 = 6.283185307179586 / 3.0;  // (2*π)/3 in Float64

// t = (νλ/4) * w^(-3/2), w = |z|/3.  w*sqrt(w) is within 1 ulp of w^(3/2)
// (sqrt and the product are each correctly rounded) and the quotient adds half
// an ulp: same accuracy class as a < 1 ulp `pow`, at a fraction of the FP64 work.
__device__ __forceinline__ double lhalf_t(double c4, double w) { return c4 / (w * sqrt(w)); }

// real part of 1 + cos(2π/3 - 2ϕ/3) with ϕ = acos(t + 0im)  (complex for t > 1:
// shiftedRootNormLhalfBox.jl:92,106): t <= 1 -> the real formula; t > 1 ->
// ϕ = -i acosh(t), cos(a + ib) = cos a cosh b - i sin a sinh b.
__device__ __forceinline__ double lhalf_one_plus_cos(double t, double cos_2pi3) {
  double re;
  if (t <= 1.0) {
    re = cos(kTwoPiOver3 - (2.0 * acos(t)) / 3.0);
  } else {
    re = cos_2pi3 * cosh((2.0 * acosh(t)) / 3.0);
  }
  return 1.0 + re;
}

// shiftedRootNormLhalf.jl:41-63
template <class R, bool PSI> struct ProxLhalf {
  using Real = R;
  static constexpr int NIN = 3, UNROLL = 2;
  static constexpr bool OUT = true, ACC = PSI;
  const R* in[NIN];  // xk, sj, q
  R fill[NIN];
  R* y;
  R nulam;      // σλ
  double p;     // 54^(1/3) (2νλ)^(2/3) / 4  (Float64)
  double c4;    // (double)(νλ/4)
  __device__ __forceinline__ R apply(const R (&x)[NIN], long long, Partial& acc) const {
    R xs = x[0] + x[1];
    R z = x[2] + xs;  // ψ.sol[i]  (:50)
    R az = jl_abs(z);
    R o;
    if ((double)az <= p) {
      o = R(0);
    } else {
      double t = lhalf_t(c4, (double)(az / R(3)));
      R coef = (R(2) * jl_sign(z)) / R(3) * az;
      o = (R)((double)coef * (1.0 + cos(kTwoPiOver3 - (2.0 * acos(t)) / 3.0)));
    }
    o = o - xs;
    if (PSI) acc.s += h_term(SPX_H_LHALF, xs + o);
    return o;
  }
};

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
template <class R, bool PSI> struct ProxL1Box {
  using Real = R;
  static constexpr int NIN = 5, UNROLL = 2;
  static constexpr bool OUT = true, ACC = PSI;
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

// shiftedNormL1Box.jl:131-225
template <class R, bool PSI> struct IproxL1Box {
  using Real = R;
  static constexpr int NIN = 6, UNROLL = 2;
  static constexpr bool OUT = true, ACC = PSI;
  const R* in[NIN];  // xk, sj, g, d, l, u
  R fill[NIN];
  R* y;
  DevSel sel;
  R lambda;
  __device__ __forceinline__ R apply(const R (&x)[NIN], long long i, Partial& acc) const {
    const R xi = x[0], si = x[1], gi = x[2], di = x[3], li = x[4], ui = x[5];
    const R eps = Eps<R>::value;
    const bool s = sel.has(i);
    const R xs = xi + si;
    const R left = li - si, right = ui - si;
    R yi;
    if (!s) {
      yi = iprox_zero(di, gi, left, right);
    } else if (jl_abs(di) <= eps) {  // :152-159
      if (jl_abs(gi) <= lambda) yi = jl_min(jl_max(left, -xs), right);
      else yi = (gi > R(0)) ? left : right;
    } else {
      const R di_2 = di / R(2);
      const R lx = li + xi, ux = ui + xi;
      const R gi2_di = gi / di_2;
      const R fi2_di = gi2_di - R(2) * xs;
      const R l2_di = lambda / di_2;
      const R val_left = lx * lx + fi2_di * lx + l2_di * jl_abs(lx);
      const R val_right = ux * ux + fi2_di * ux + l2_di * jl_abs(ux);
      if (di > eps) {  // :161-198
        R val_min = jl_min(val_left, val_right);
        yi = (val_left < val_right) ? left : right;
        const R a1 = (-(gi + lambda)) / di;
        const R a2 = (lambda - gi) / di;
        const bool in1 = (left <= a1) && (a1 <= right);
        const bool in2 = (left <= a2) && (a2 <= right);
        if (lx >= R(0)) {
          if (in1) yi = a1;
        } else if (R(0) >= ux) {
          if (in2) yi = a2;
        } else {
          if (in1) {
            R v1 = xs + a1;
            R val1 = v1 * v1 + fi2_di * v1 + l2_di * jl_abs(v1);
            if (val1 < val_min) yi = a1;
            val_min = jl_min(val1, val_min);
          }
          if (in2) {
            R v2 = xs + a2;
            R val2 = v2 * v2 + fi2_di * v2 + l2_di * jl_abs(v2);
            if (val2 < val_min) yi = a2;
            val_min = jl_min(val2, val_min);
          }
          if (R(0) < val_min) yi = -xs;
        }
      } else {  // :200-218
        R val_max = jl_max(val_left, val_right);
        yi = (val_left > val_right) ? left : right;
        if ((li <= -xi) && (-xi <= ui)) {
          if (R(0) > val_max) yi = -xs;
        }
      }
    }
    if (PSI) BoxPsi<R>{SPX_H_L1}.add(acc, s, xi, si, yi, li, ui);
    return yi;
  }
};

// shiftedNormL0Box.jl:89-131
template <class R, bool PSI> struct ProxL0Box {
  using Real = R;
  static constexpr int NIN = 5, UNROLL = 2;
  static constexpr bool OUT = true, ACC = PSI;
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

// shiftedNormL0Box.jl:137-231
template <class R, bool PSI> struct IproxL0Box {
  using Real = R;
  static constexpr int NIN = 6, UNROLL = 2;
  static constexpr bool OUT = true, ACC = PSI;
  const R* in[NIN];  // xk, sj, g, d, l, u
  R fill[NIN];
  R* y;
  DevSel sel;
  R lambda;
  __device__ __forceinline__ R apply(const R (&x)[NIN], long long i, Partial& acc) const {
    const R xi = x[0], si = x[1], gi = x[2], di = x[3], li = x[4], ui = x[5];
    const R eps = Eps<R>::value;
    const bool s = sel.has(i);
    const R xs = xi + si;
    const R left = li - si, right = ui - si;
    const bool zero_in = (li <= -xi) && (-xi <= ui);
    R yi;
    if (!s) {
      yi = iprox_zero(di, gi, left, right);
    } else if (jl_abs(di) < eps) {  // :155-177
      if (gi == R(0)) {
        yi = zero_in ? -xs : R(0);
      } else {
        // gi > 0 -> left edge, gi < 0 -> right edge; a NaN g leaves y[i] untouched in the
        // reference (no branch taken): NaN is written here (y is write-only)
        R val_min = gi - gi;  // 0, or NaN for NaN g
        yi = val_min / val_min;  // NaN placeholder, overwritten below for finite g
        if (gi > R(0)) {
          val_min = gi * left + ((xi == -li) ? R(0) : lambda);
          yi = left;
        } else if (gi < R(0)) {
          val_min = gi * right + ((xi == -ui) ? R(0) : lambda);
          yi = right;
        }
        if (zero_in) {
          R val_0 = (-gi) * xs;
          if (val_0 < val_min) yi = -xs;
        }
      }
    } else {  // :179-224
      const R di_2 = di / R(2);
      const R lx = li + xi, ux = ui + xi;
      const R gi2_di = gi / di_2;
      const R fi2_di = gi2_di - R(2) * xs;
      const R l2_di = lambda / di_2;
      const R val_left = (lx == R(0)) ? R(0) : (lx * lx + fi2_di * lx + l2_di);
      const R val_right = (ux == R(0)) ? R(0) : (ux * ux + fi2_di * ux + l2_di);
      if (di >= eps) {  // :189-209
        const R aq_y = (-gi) / di;
        const R aq_v = aq_y + xs;
        R val_min;
        if ((lx <= aq_v) && (aq_v <= ux)) {
          val_min = (aq_v == R(0)) ? -(aq_v * aq_v) : (-(aq_v * aq_v) + l2_di);
          yi = aq_y;
        } else {
          yi = (val_left < val_right) ? left : right;
          val_min = jl_min(val_left, val_right);
        }
        if (zero_in && (R(0) < val_min)) yi = -xs;
      } else {  // :211-223
        yi = (val_left > val_right) ? left : right;
        R val_max = jl_max(val_left, val_right);
        if (zero_in && (R(0) > val_max)) yi = -xs;
      }
    }
    if (PSI) BoxPsi<R>{SPX_H_L0}.add(acc, s, xi, si, yi, li, ui);
    return yi;
  }
};

// Base.isless / isgreater on Float64 (findmin's ordering: NaN counts as minimal,
// first minimal index wins) -- shiftedRootNormLhalfBox.jl:108
__device__ __forceinline__ bool jl_isless(double a, double b) {
  if (a != a) return false;
  if (b != b) return true;
  if (a == b) return sgnbit(a) && !sgnbit(b);
  return a < b;
}
__device__ __forceinline__ bool jl_isgreater(double x, double y) {
  return (x != x || y != y) ? jl_isless(x, y) : jl_isless(y, x);
}

// shiftedRootNormLhalfBox.jl:86-120
template <class R, bool PSI> struct ProxLhalfBox {
  using Real = R;
  static constexpr int NIN = 5, UNROLL = 2;
  static constexpr bool OUT = true, ACC = PSI;
  const R* in[NIN];  // xk, sj, q, l, u
  R fill[NIN];
  R* y;
  DevSel sel;
  R lambda, sigma;
  double c4;        // (double)(σλ/4), σλ/4 evaluated in R
  double cos_2pi3;  // cos((2π)/3) in Float64
  __device__ __forceinline__ R apply(const R (&x)[NIN], long long i, Partial& acc) const {
    const R xi = x[0], si = x[1], qi = x[2], li = x[3], ui = x[4];
    const bool s = sel.has(i);
    R o;
    if (!s) {
      o = prox_zero(qi, li - si, ui - si);
    } else {
      const R xs = xi + si;  // ψ.sol[i]  (:94)
      const R xsq = xs + qi;
      const R axsq = jl_abs(xsq);
      const double t = lhalf_t(c4, (double)(axsq / R(3)));
      const R coef = (R(2) * jl_sign(xsq)) / R(3) * axsq;
      const double val = (double)coef * lhalf_one_plus_cos(t, cos_2pi3);
      // RNorm(tt) = (tt - q)^2 / 2 / σ + λ sqrt(|tt + xs|)   (:95)
      const R left = li - si, right = ui - si, mxs = -xs;
      auto rnorm = [&](R tt) -> double {
        R dq = tt - qi;
        return (double)((dq * dq) / R(2) / sigma + lambda * sqrt(jl_abs(tt + xs)));
      };
      double c0 = rnorm(left);
      double c1 = rnorm(right);
      const double inf = __longlong_as_double(0x7ff0000000000000ll);
      double c2 = ((li <= -xi) && (-xi <= ui)) ? rnorm(mxs) : inf;
      const double vmx = val - (double)xi;
      const double cand4 = val - (double)xs;  // Float64 (val is Float64)
      double c3 = inf;
      if (((double)li <= vmx) && (vmx <= (double)ui)) {
        double dq = cand4 - (double)qi;
        c3 = (dq * dq) / 2.0 / (double)sigma + (double)lambda * sqrt(fabs(cand4 + (double)xs));
      }
      int a = 0;
      double fm = c0;
      if (jl_isgreater(fm, c1)) { fm = c1; a = 1; }
      if (jl_isgreater(fm, c2)) { fm = c2; a = 2; }
      if (jl_isgreater(fm, c3)) { fm = c3; a = 3; }
      o = a == 0 ? left : (a == 1 ? right : (a == 2 ? mxs : (R)cand4));
    }
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
  static constexpr int NIN = 5, UNROLL = 2;
  static constexpr bool OUT = false, ACC = true;
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
