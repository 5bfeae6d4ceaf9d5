// ============================================================================
// oracle/shifted_prox_oracle.cpp -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
//
// CPU restatement of the shifted prox hot path of ShiftedProximalOperators.jl
// v0.2.2 (pure Julia; `julia` is not installed in this image, so the reference
// itself cannot be executed here).  Every function cites the reference
// file:line it follows.  Only tests/, __graft_entry__.smoke() and bench.py's
// cpu_baseline / --impl reference legs may load this library; the product
// (libshiftedprox.so) never does.
//
// Pinning: checked by tests/test_oracle_golden.py against every known-answer
// vector the reference's own tests hold for this path (test/runtests.jl:113-126,
// 449-494, 587-606, 694-705, 814-843; test/testsbox.jl:19-22,47-49,73-76,
// 115-179,207-271; test/partial_prox.jl:33-72).
// PARITY UNPINNED (no reference test fixes a value; the cited source text is
// the only specification): ShiftedNormL1/L0 prox!/iprox!, ShiftedRootNormLhalf
// prox!, ShiftedIndBallL0(BInf) prox! incl. the tie rule, all Float32 values,
// and the iterates of the third-party root finders (Roots.jl ^1.0, not vendored
// under /root/reference) beyond the reference's own `≈` criterion.
//
// Third-party semantics restated here (sources absent from /root/reference):
//   ProximalOperators.jl 0.15 (NormL1/NormL0/NormL2/IndBallL0/IndBallL2/
//   IndBallLinf value functors), Roots.jl ^1.0 (find_zero / fzero),
//   Base.sortperm!, LinearAlgebra.norm, Base.min/max/sign/findmin.
//
// Build: g++ -O2 -std=c++17 -ffp-contract=off -fno-fast-math -shared -fPIC
// (Julia never contracts a*b+c and never reassociates: no @fastmath/@simd/
// muladd anywhere in the reference).
// ============================================================================
#include <algorithm>
#include <cmath>
#include <complex>
#include <cstdint>
#include <cstring>
#include <limits>
#include <numeric>
#include <vector>

namespace {

// ---------------------------------------------------------------- Base.* --
// Julia Base.min/max for IEEE floats: NaN-propagating, -0.0 < +0.0.
template <class R> inline R jl_min(R x, R y) {
  R diff = x - y;
  R arg = std::signbit(diff) ? x : y;
  return (std::isnan(x) || std::isnan(y)) ? diff : arg;
}
template <class R> inline R jl_max(R x, R y) {
  R diff = x - y;
  R arg = std::signbit(diff) ? y : x;
  return (std::isnan(x) || std::isnan(y)) ? diff : arg;
}
// Base.sign: sign(±0.0) = ±0.0, sign(NaN) = NaN.
template <class R> inline R jl_sign(R x) { return x > 0 ? R(1) : (x < 0 ? R(-1) : x); }
// Base.isless on floats: total order, -0.0 < +0.0, NaN largest.
inline bool jl_isless(double a, double b) {
  if (std::isnan(a)) return false;
  if (std::isnan(b)) return true;
  if (a == b) return std::signbit(a) && !std::signbit(b);
  return a < b;
}
// Base.isgreater (used by findmin's reducer): NaN counts as smallest.
inline bool jl_isgreater(double x, double y) {
  return (std::isnan(x) || std::isnan(y)) ? jl_isless(x, y) : jl_isless(y, x);
}

// `selected::AbstractArray{<:Integer}` (shiftedNormL1Box.jl:10,19).  The
// reference tests membership with `i ∈ ψ.selected` (:106) and gathers
// `x[ψ.selected]` in ψ(y) (:71), so duplicates count twice in the value.
// kind 0 = every index; kind 1 = explicit 0-based list (any order, duplicates).
struct Sel {
  int kind;
  const int64_t* list;
  int64_t nlist;
  std::vector<uint8_t> member;  // membership table built once per call
  Sel(int k, const int64_t* l, int64_t nl, int64_t n) : kind(k), list(l), nlist(nl) {
    if (kind == 1) {
      member.assign((size_t)n, 0);
      for (int64_t j = 0; j < nl; ++j)
        if (l[j] >= 0 && l[j] < n) member[(size_t)l[j]] = 1;
    }
  }
  inline bool has(int64_t i) const { return kind == 0 ? true : member[(size_t)i] != 0; }
};

template <class R> struct Bound {  // `l`/`u` may be a scalar or a vector (V3, V4 untyped)
  const R* vec;
  R val;
  inline R at(int64_t i) const { return vec ? vec[i] : val; }
};

// ShiftedProximalOperators.jl:203
template <class R> inline R prox_zero(R q, R l, R u) { return jl_min(jl_max(q, l), u); }

// ShiftedProximalOperators.jl:217-236
template <class R> inline R iprox_zero(R d, R g, R l, R u) {
  const R eps = std::numeric_limits<R>::epsilon();
  if (d > eps) {
    R argmin_quad = (-g) / d;
    return jl_min(jl_max(argmin_quad, l), u);
  } else if (d < -eps) {
    R d_2 = d / R(2);
    R val_l = d_2 * (l * l) + g * l;
    R val_u = d_2 * (u * u) + g * u;
    return (val_l < val_u) ? l : u;
  } else {
    if (g > R(0)) return l;
    if (g < R(0)) return u;
    return R(0);
  }
}

// ------------------------------------------------ ShiftedNormL1 (a2, a3) --
// shiftedNormL1.jl:40-54 (two passes: broadcast :47, loop :49-51)
template <class R>
void prox_l1(int64_t n, R* y, const R* xk, const R* sj, const R* q, R lambda, R sigma) {
  for (int64_t i = 0; i < n; ++i) y[i] = (-xk[i]) - sj[i];
  for (int64_t i = 0; i < n; ++i) {
    R a = lambda * sigma;
    y[i] = jl_min(jl_max(y[i], q[i] - a), q[i] + a);
  }
}
// shiftedNormL1.jl:60-75.  Returns the index of the first d[i] <= 0 (the
// reference throws AssertionError there, :70) or -1.
template <class R>
int64_t iprox_l1(int64_t n, R* y, const R* xk, const R* sj, const R* g, const R* d, R lambda) {
  for (int64_t i = 0; i < n; ++i) y[i] = (-xk[i]) - sj[i];
  for (int64_t i = 0; i < n; ++i) {
    if (!(d[i] > 0)) return i;
    R c = (-g[i]) / d[i];
    R w = lambda / d[i];
    y[i] = jl_min(jl_max(y[i], c - w), c + w);
  }
  return -1;
}

// ------------------------------------------------ ShiftedNormL0 (a4, a5) --
// shiftedNormL0.jl:38-55
template <class R>
void prox_l0(int64_t n, R* y, const R* xk, const R* sj, const R* q, R lambda, R sigma) {
  R c = std::sqrt(R(2) * lambda * sigma);
  for (int64_t i = 0; i < n; ++i) {
    R xps = xk[i] + sj[i];
    R qi = q[i];
    y[i] = (std::fabs(xps + qi) <= c) ? -xps : qi;
  }
}
// shiftedNormL0.jl:61-80
template <class R>
int64_t iprox_l0(int64_t n, R* y, const R* xk, const R* sj, const R* g, const R* d, R lambda) {
  for (int64_t i = 0; i < n; ++i) {
    R di = d[i];
    if (!(di > 0)) return i;
    R ci = std::sqrt(R(2) * lambda * di);
    R xps = xk[i] + sj[i];
    R gi = g[i];
    y[i] = (std::fabs(di * xps - gi) <= ci) ? -xps : (-gi) / di;
  }
  return -1;
}

// ------------------------------------------- RootNormLhalf closed form ----
// Threshold p of shiftedRootNormLhalf.jl:49 / rootNormLhalf.jl:40:
// `54^(1/3) * (2νλ)^(2/3) / 4` -- Float64 whatever R is (Int^Float64 and
// R^Float64 both promote to Float64).
template <class R> inline double lhalf_threshold(R nulam) {
  return std::pow(54.0, 1.0 / 3.0) * std::pow((double)(R(2) * nulam), 2.0 / 3.0) / 4.0;
}
// Real branch of the closed form, shiftedRootNormLhalf.jl:48,57 /
// rootNormLhalf.jl:38,45: `2*sign(z)/3*|z|*(1+cos(2π/3 - 2ϕ(z)/3))`,
// ϕ(z) = acos(νλ/4 * (|z|/3)^(-3/2)).  The first three factors stay in R,
// the power/acos/cos run in Float64.
template <class R> inline double lhalf_closed_form(R z, R nulam) {
  R az = std::fabs(z);
  double t = (double)(nulam / R(4)) * std::pow((double)(az / R(3)), -1.5);
  double phi = std::acos(t);
  const double two_pi_3 = 6.283185307179586 / 3.0;  // (2*π)/3 in Float64
  R coef = (R(2) * jl_sign(z)) / R(3) * az;
  return (double)coef * (1.0 + std::cos(two_pi_3 - (2.0 * phi) / 3.0));
}

// rootNormLhalf.jl:31-51 (unshifted; returns λ·Σ√|y_i| over the kept entries)
template <class R> double prox_rootlhalf_unshifted(int64_t n, R* y, const R* x, R lambda, R gamma) {
  R gl = gamma * lambda;
  double thr = lhalf_threshold(gl);
  R ysum = 0;
  for (int64_t i = 0; i < n; ++i) {
    if ((double)std::fabs(x[i]) <= thr) {
      y[i] = 0;
    } else {
      y[i] = (R)lhalf_closed_form(x[i], gl);
      ysum += std::sqrt(std::fabs(y[i]));
    }
  }
  return (double)(lambda * ysum);
}

// ------------------------------------------ ShiftedRootNormLhalf (a6) -----
// shiftedRootNormLhalf.jl:41-63; `sol` is ψ.sol (:50).
template <class R>
void prox_lhalf(int64_t n, R* y, R* sol, const R* xk, const R* sj, const R* q, R lambda, R sigma) {
  R nulam = sigma * lambda;
  double p = lhalf_threshold(nulam);
  for (int64_t i = 0; i < n; ++i) sol[i] = q[i] + (xk[i] + sj[i]);
  for (int64_t i = 0; i < n; ++i) {
    R aqi = std::fabs(sol[i]);
    if ((double)aqi <= p) y[i] = 0;
    else y[i] = (R)lhalf_closed_form(sol[i], nulam);
    y[i] -= (xk[i] + sj[i]);
  }
}

// --------------------------------------------- ShiftedNormL1Box (a8, a9) --
// shiftedNormL1Box.jl:89-125
template <class R>
void prox_l1box(int64_t n, R* y, const R* xk, const R* sj, const R* q, Bound<R> l, Bound<R> u,
                const Sel& sel, R lambda, R sigma) {
  R sl = sigma * lambda;
  for (int64_t i = 0; i < n; ++i) {
    R li = l.at(i), ui = u.at(i), qi = q[i], si = sj[i];
    if (sel.has(i)) {
      R xs = xk[i] + si;
      R xsq = xs + qi;
      R yi = (xsq <= -sl) ? qi + sl : (xsq >= sl) ? qi - sl : -xs;
      y[i] = jl_min(jl_max(yi, li - si), ui - si);
    } else {
      y[i] = prox_zero(qi, li - si, ui - si);
    }
  }
}
// shiftedNormL1Box.jl:131-225
template <class R>
void iprox_l1box(int64_t n, R* y, const R* xk, const R* sj, const R* g, const R* d, Bound<R> l,
                 Bound<R> u, const Sel& sel, R lambda) {
  const R eps = std::numeric_limits<R>::epsilon();
  for (int64_t i = 0; i < n; ++i) {
    R li = l.at(i), ui = u.at(i), di = d[i], gi = g[i], si = sj[i], xi = xk[i];
    R xs = xi + si;
    if (!sel.has(i)) {
      y[i] = iprox_zero(di, gi, li - si, ui - si);
      continue;
    }
    R left = li - si, right = ui - si;
    R yi;
    if (std::fabs(di) <= eps) {  // :152-159
      if (std::fabs(gi) <= lambda) yi = jl_min(jl_max(left, -xs), right);
      else yi = (gi > 0) ? left : right;
    } else if (di > eps) {  // :161-198
      R di_2 = di / R(2);
      R lx = li + xi, ux = ui + xi;
      R gi2_di = gi / di_2;
      R fi2_di = gi2_di - R(2) * xs;
      R l2_di = lambda / di_2;
      R val_left = lx * lx + fi2_di * lx + l2_di * std::fabs(lx);
      R val_right = ux * ux + fi2_di * ux + l2_di * std::fabs(ux);
      R val_min = jl_min(val_left, val_right);
      yi = (val_left < val_right) ? left : right;
      if (lx >= R(0)) {
        R a = (-(gi + lambda)) / di;
        if (left <= a && a <= right) yi = a;
      } else if (R(0) >= ux) {
        R a = (lambda - gi) / di;
        if (left <= a && a <= right) yi = a;
      } else {
        R a1 = (-(gi + lambda)) / di;
        R a2 = (lambda - gi) / di;
        if (left <= a1 && a1 <= right) {
          R v1 = xs + a1;
          R val1 = v1 * v1 + fi2_di * v1 + l2_di * std::fabs(v1);
          if (val1 < val_min) yi = a1;
          val_min = jl_min(val1, val_min);
        }
        if (left <= a2 && a2 <= right) {
          R v2 = xs + a2;
          R val2 = v2 * v2 + fi2_di * v2 + l2_di * std::fabs(v2);
          if (val2 < val_min) yi = a2;
          val_min = jl_min(val2, val_min);
        }
        if (R(0) < val_min) yi = -xs;
      }
    } else {  // :200-218
      R di_2 = di / R(2);
      R gi2_di = gi / di_2;
      R fi2_di = gi2_di - R(2) * xs;
      R l2_di = lambda / di_2;
      R lx = li + xi, ux = ui + xi;
      R val_left = lx * lx + fi2_di * lx + l2_di * std::fabs(lx);
      R val_right = ux * ux + fi2_di * ux + l2_di * std::fabs(ux);
      R val_max = jl_max(val_left, val_right);
      yi = (val_left > val_right) ? left : right;
      if (li <= -xi && -xi <= ui) {
        if (R(0) > val_max) yi = -xs;
      }
    }
    y[i] = yi;
  }
}

// -------------------------------------------- ShiftedNormL0Box (a10, a11) --
// shiftedNormL0Box.jl:89-131
template <class R>
void prox_l0box(int64_t n, R* y, const R* xk, const R* sj, const R* q, Bound<R> l, Bound<R> u,
                const Sel& sel, R lambda, R sigma) {
  R c = R(2) * lambda * sigma;
  for (int64_t i = 0; i < n; ++i) {
    R li = l.at(i), ui = u.at(i), qi = q[i], si = sj[i];
    R sq = si + qi;
    if (sel.has(i)) {
      R xi = xk[i];
      R xs = xi + si;
      R xsq = xs + qi;
      R dl = li - sq, du = ui - sq;
      R val_left = dl * dl + (xi == -li ? R(0) : c);
      R val_right = du * du + (xi == -ui ? R(0) : c);
      R yi = (val_left < val_right) ? (li - si) : (ui - si);
      R val_min = jl_min(val_left, val_right);
      if (li <= -xi && -xi <= ui) {
        R val_0 = xsq * xsq;
        if (val_0 < val_min) yi = -xs;
        val_min = jl_min(val_0, val_min);
      }
      if (li <= sq && sq <= ui) {
        R val_xsq = (xsq == R(0)) ? R(0) : c;
        if (val_xsq < val_min) yi = qi;
      }
      y[i] = yi;
    } else {
      y[i] = prox_zero(qi, li - si, ui - si);
    }
  }
}
// shiftedNormL0Box.jl:137-231
template <class R>
void iprox_l0box(int64_t n, R* y, const R* xk, const R* sj, const R* g, const R* d, Bound<R> l,
                 Bound<R> u, const Sel& sel, R lambda) {
  const R eps = std::numeric_limits<R>::epsilon();
  for (int64_t i = 0; i < n; ++i) {
    R li = l.at(i), ui = u.at(i), di = d[i], gi = g[i], si = sj[i], xi = xk[i];
    R xs = xi + si;
    if (!sel.has(i)) {
      y[i] = iprox_zero(di, gi, li - si, ui - si);
      continue;
    }
    R yi = y[i];
    const bool zero_in = (li <= -xi && -xi <= ui);
    if (std::fabs(di) < eps) {  // :155-177
      if (gi == R(0)) {
        yi = zero_in ? -xs : R(0);
      } else {
        R val_min = std::numeric_limits<R>::quiet_NaN();
        if (gi > R(0)) {
          R left = li - si;
          val_min = gi * left + (xi == -li ? R(0) : lambda);
          yi = left;
        } else if (gi < R(0)) {
          R right = ui - si;
          val_min = gi * right + (xi == -ui ? R(0) : lambda);
          yi = right;
        }
        if (zero_in) {
          R val_0 = (-gi) * xs;
          if (val_0 < val_min) yi = -xs;
        }
      }
    } else {  // :179-224
      R di_2 = di / R(2);
      R left = li - si, right = ui - si;
      R lx = li + xi, ux = ui + xi;
      R gi2_di = gi / di_2;
      R fi2_di = gi2_di - R(2) * xs;
      R l2_di = lambda / di_2;
      if (di >= eps) {  // :189-209
        R aq_y = (-gi) / di;
        R aq_v = aq_y + xs;
        R val_min;
        if (lx <= aq_v && aq_v <= ux) {
          val_min = (aq_v == R(0)) ? -(aq_v * aq_v) : (-(aq_v * aq_v) + l2_di);
          yi = aq_y;
        } else {
          R val_left = (lx == R(0)) ? R(0) : (lx * lx + fi2_di * lx + l2_di);
          R val_right = (ux == R(0)) ? R(0) : (ux * ux + fi2_di * ux + l2_di);
          yi = (val_left < val_right) ? left : right;
          val_min = jl_min(val_left, val_right);
        }
        if (zero_in) {
          if (R(0) < val_min) yi = -xs;
        }
      } else {  // :211-223
        R val_left = (lx == R(0)) ? R(0) : (lx * lx + fi2_di * lx + l2_di);
        R val_right = (ux == R(0)) ? R(0) : (ux * ux + fi2_di * ux + l2_di);
        yi = (val_left > val_right) ? left : right;
        R val_max = jl_max(val_left, val_right);
        if (zero_in) {
          if (R(0) > val_max) yi = -xs;
        }
      }
    }
    y[i] = yi;
  }
}

// ------------------------------------- ShiftedRootNormLhalfBox (a12) ------
// shiftedRootNormLhalfBox.jl:86-120.  `val` uses complex acos/cos (:92,106):
// Julia's acos(t+0im) (Kahan) and cos(::Complex) = cos(a)cosh(b) - i sin(a)sinh(b);
// restated with std::complex<double> (glibc cacos/ccos), real part taken.
template <class R>
void prox_lhalfbox(int64_t n, R* y, R* sol, const R* xk, const R* sj, const R* q, Bound<R> l,
                   Bound<R> u, const Sel& sel, R lambda, R sigma, int8_t* pick = nullptr,
                   double* obj = nullptr) {
  // pick / obj (tests only): index of the candidate `findmin` chose (-1: not selected) and the
  // four objectives (c[0..3]) plus candidate 4 itself (obj[5 i + 4]) for every element, so a parity test can tell a true tie from a wrong pick
  for (int64_t i = 0; i < n; ++i) sol[i] = xk[i] + sj[i];  // :94
  const double two_pi_3 = 6.283185307179586 / 3.0;
  const double inf = std::numeric_limits<double>::infinity();
  for (int64_t i = 0; i < n; ++i) {
    R li = l.at(i), ui = u.at(i), xi = xk[i], si = sj[i], qi = q[i];
    if (!sel.has(i)) {
      y[i] = prox_zero(qi, li - si, ui - si);
      if (pick) pick[i] = -1;
      continue;
    }
    R xs = sol[i];
    R xsq = xs + qi;
    R axsq = std::fabs(xsq);
    // ϕ(z) = acos(σ*λ/4 * (|z|/3)^(-3/2) + 0im)
    double t = (double)(sigma * lambda / R(4)) * std::pow((double)(axsq / R(3)), -1.5);
    std::complex<double> phi = std::acos(std::complex<double>(t, 0.0));
    std::complex<double> arg(two_pi_3 - (2.0 * phi.real()) / 3.0, -((2.0 * phi.imag()) / 3.0));
    std::complex<double> cs = std::cos(arg);
    R coef = (R(2) * jl_sign(xsq)) / R(3) * axsq;
    double val = (double)coef * (1.0 + cs.real());
    // RNorm(tt, i) = (tt - q[i])^2 / 2 / σ + λ * sqrt(abs(tt + sol[i]))   (:95)
    auto rnorm_R = [&](R tt) -> double {
      R dq = tt - qi;
      return (double)((dq * dq) / R(2) / sigma + lambda * std::sqrt(std::fabs(tt + xs)));
    };
    auto rnorm_D = [&](double tt) -> double {  // candidate 4: tt is Float64 (val - xs)
      double dq = tt - (double)qi;
      return (dq * dq) / 2.0 / (double)sigma + (double)lambda * std::sqrt(std::fabs(tt + (double)xs));
    };
    double c[4];
    c[0] = rnorm_R(li - si);
    c[1] = rnorm_R(ui - si);
    c[2] = (li <= -xi && -xi <= ui) ? rnorm_R(-xs) : inf;
    double vmx = val - (double)xi;
    c[3] = ((double)li <= vmx && vmx <= (double)ui) ? rnorm_D(val - (double)xs) : inf;
    int a = 0;  // findmin: first minimal index, NaN counts as minimal
    double fm = c[0];
    for (int k = 1; k < 4; ++k)
      if (jl_isgreater(fm, c[k])) { fm = c[k]; a = k; }
    y[i] = a == 0 ? (li - si) : a == 1 ? (ui - si) : a == 2 ? -xs : (R)(val - (double)xs);
    if (pick) pick[i] = (int8_t)a;
    if (obj) {
      for (int k = 0; k < 4; ++k) obj[5 * i + k] = c[k];
      obj[5 * i + 4] = val - (double)xs;  // candidate 4 itself (Float64, before the store rounds it to R)
    }
  }
}

// ----------------------------------------------------- norms / sums -------
// LinearAlgebra.norm(x) (p = 2).  The reference leaves the summation order
// unspecified (generic sequential Float64 accumulation below 32 elements,
// OpenBLAS nrm2 above); restated as a sequential sum of squares in extended
// precision, i.e. the correctly rounded value up to ~1 ulp.
template <class R> inline R norm2(const R* x, int64_t n) {
  long double s = 0;
  for (int64_t i = 0; i < n; ++i) s += (long double)x[i] * (long double)x[i];
  return (R)std::sqrt(s);
}

// ------------------------------------------------- ShiftedNormL1B2 (a13) --
// shiftedNormL1B2.jl:47-64.  Roots.find_zero(froot, Δ) (Order0) is third party
// and not vendored; restated as: bracket [Δ, ηhi] with froot(Δ) <= 0 <= froot(ηhi)
// then bisection on the float lattice of R down to adjacent floats -- any
// solver converging to the root gives the same y up to a few ulp (the reference
// test only asks `≈`, runtests.jl:493).  Returns the number of froot evaluations.
template <class R>
int prox_l1b2(int64_t n, R* y, const R* xk, const R* sj, const R* q, R lambda, R sigma, R delta,
              R chi_lambda) {
  R ls = lambda * sigma;
  std::vector<R> w((size_t)n);
  auto projB = [&](R scale, bool scaled, R* out) {
    for (int64_t i = 0; i < n; ++i) {
      R z = scaled ? (-xk[i]) * scale : -xk[i];
      R lo = (sj[i] + q[i]) - ls, hi = (sj[i] + q[i]) + ls;
      out[i] = jl_min(jl_max(z, lo), hi);
    }
  };
  auto chi = [&](const R* v) -> R { return chi_lambda * norm2(v, n); };
  int evals = 0;
  projB(R(0), false, y);
  if (delta <= chi(y)) {
    auto froot = [&](R eta) -> R {
      ++evals;
      projB(eta / delta, true, w.data());
      return eta - chi(w.data());
    };
    R a = delta, fa = froot(a);
    R eta = a;
    if (fa != R(0)) {
      R b = jl_max(R(2) * a, a + R(1));
      R fb = froot(b);
      while (fb < R(0) && std::isfinite(b)) { a = b; fa = fb; b = R(2) * b; fb = froot(b); }
      // bisection until a and b are adjacent floats (or an exact zero is hit)
      eta = b;
      if (fb != R(0)) {
        while (true) {
          R m = a + (b - a) / R(2);
          if (!(a < m && m < b)) break;
          R fm = froot(m);
          if (fm == R(0)) { a = b = m; break; }
          if (fm < R(0)) { a = m; fa = fm; } else { b = m; fb = fm; }
        }
        eta = (std::fabs(fa) <= std::fabs(fb)) ? a : b;
      }
    }
    projB(eta / delta, true, w.data());
    R post = delta / eta;
    for (int64_t i = 0; i < n; ++i) y[i] = w[i] * post;
  }
  for (int64_t i = 0; i < n; ++i) y[i] -= sj[i];
  return evals;
}

// ---------------------------------------------- ShiftedGroupNormL2 (a14) --
// shiftedGroupNormL2.jl:52-79.  Groups are contiguous index ranges given as
// CSR offsets (ngroups+1 entries), weights lambda[g]  (groupNormL2.jl:15-31).
template <class R>
void prox_groupl2(int64_t n, R* y, R* sol, const R* xk, const R* sj, const R* q, int64_t ngroups,
                  const int64_t* offs, const R* lambda, R sigma) {
  for (int64_t i = 0; i < n; ++i) sol[i] = (q[i] + xk[i]) + sj[i];  // :65
  for (int64_t g = 0; g < ngroups; ++g) {
    int64_t b = offs[g], e = offs[g + 1];
    R snorm = norm2(sol + b, e - b);
    if (snorm == R(0)) {
      for (int64_t i = b; i < e; ++i) y[i] = 0;
    } else {
      R alpha = jl_max(R(1) - sigma * lambda[g] / snorm, R(0));
      for (int64_t i = b; i < e; ++i) y[i] = alpha * sol[i];
    }
  }
  for (int64_t i = 0; i < n; ++i) y[i] -= (xk[i] + sj[i]);  // :77
}
// groupNormL2.jl:41-58 (unshifted; returns Σ λ_g‖x_g‖ over nonzero groups)
template <class R>
double prox_groupl2_unshifted(int64_t n, R* y, const R* x, int64_t ngroups, const int64_t* offs,
                              const R* lambda, R gamma) {
  (void)n;
  R ysum = 0;
  for (int64_t g = 0; g < ngroups; ++g) {
    int64_t b = offs[g], e = offs[g + 1];
    R yt = norm2(x + b, e - b);
    if (yt == R(0)) {
      for (int64_t i = b; i < e; ++i) y[i] = 0;
    } else {
      R a = jl_max(R(1) - gamma * lambda[g] / yt, R(0));
      for (int64_t i = b; i < e; ++i) y[i] = a * x[i];
      ysum += lambda[g] * yt;
    }
  }
  return (double)ysum;
}

// ------------------------------------------ ShiftedGroupNormL2Binf (a15) --
// shiftedGroupNormL2Binf.jl:67-119.  Roots.fzero(froot, lmin, lmax) =
// find_zero(f, (a,b), Bisection()): bisection on the float lattice until the
// bracket is two adjacent floats or an exact zero (third party, restated).
template <class R>
void prox_groupl2binf(int64_t n, R* y, R* sol, const R* xk, const R* sj, const R* q,
                      int64_t ngroups, const int64_t* offs, const R* lambda, R sigma, R delta,
                      double* nroot_out = nullptr, int8_t* zero_out_g = nullptr, int64_t ulp_shift = 0) {
  // nroot_out / zero_out_g / ulp_shift (tests only): the root each group's bisection ended on, whether the
  // group was zeroed, and a shift of that root by whole ulps before the final formula -- the sensitivity of y
  // to the last bit of the root, i.e. the conditioning a parity tolerance has to allow for
  const R eps = std::numeric_limits<R>::epsilon();
  for (int64_t i = 0; i < n; ++i) sol[i] = (q[i] + xk[i]) + sj[i];  // :80
  std::vector<R> tmp;
  auto softthres = [](R x, R a) -> R { return jl_sign(x) * jl_max(R(0), std::fabs(x) - a); };
  for (int64_t g = 0; g < ngroups; ++g) {
    int64_t b = offs[g], e = offs[g + 1], m = e - b;
    tmp.resize((size_t)m);
    const R* so = sol + b;
    const R* xg = xk + b;
    R lam = lambda[g];
    R sl = lam * sigma;
    auto cstep = [&](R nn) -> R { return nn / (sigma * (nn - sl)); };
    auto froot = [&](R nn) -> R {  // :87-93
      R c = cstep(nn);
      for (int64_t i = 0; i < m; ++i)
        tmp[i] = sigma * softthres(so[i] / sigma - c * xg[i], delta * c) - so[i];
      return nn - norm2(tmp.data(), m);
    };
    R lmin = sl * (R(1) + eps);
    R fl = froot(lmin);
    R ansatz = lmin + R(1);
    R step = ansatz / (sigma * (ansatz - sl));
    for (int64_t i = 0; i < m; ++i) tmp[i] = softthres(so[i] / sigma - step * xg[i], delta * step);
    R zlmax = norm2(tmp.data(), m);
    // abs((ϵ - 1)/ϵ + 1) with ϵ = 1 is exactly 1   (:100)
    R lmax = norm2(so, m) + sigma * (zlmax + R(1) * lam * norm2(xg, m));
    R fm = froot(lmax);
    bool zero_out = false;
    R nroot = 0;
    if (fl * fm > R(0)) {
      zero_out = true;
    } else {
      R a = lmin, fa = fl, bb = lmax, fb = fm;
      if (fa == R(0)) nroot = a;
      else if (fb == R(0)) nroot = bb;
      else {
        while (true) {
          R mid = a + (bb - a) / R(2);
          if (!(a < mid && mid < bb)) break;
          R fmid = froot(mid);
          if (fmid == R(0)) { a = bb = mid; fa = fb = 0; break; }
          if ((fmid < R(0)) == (fa < R(0))) { a = mid; fa = fmid; } else { bb = mid; fb = fmid; }
        }
        nroot = (std::fabs(fa) <= std::fabs(fb)) ? a : bb;
      }
      for (int64_t k = 0; k < ulp_shift; ++k) nroot = std::nextafter(nroot, std::numeric_limits<R>::infinity());
      for (int64_t k = 0; k > ulp_shift; --k) nroot = std::nextafter(nroot, -std::numeric_limits<R>::infinity());
      step = cstep(nroot);
      if (std::fabs(nroot - sl) == R(0)) zero_out = true;  // `abs(n - σλ) ≈ 0`  (:107)
    }
    if (nroot_out) nroot_out[g] = (fl * fm > R(0)) ? std::numeric_limits<double>::quiet_NaN() : (double)nroot;
    if (zero_out_g) zero_out_g[g] = zero_out ? 1 : 0;
    if (zero_out) {
      for (int64_t i = b; i < e; ++i) y[i] = 0;
    } else {
      for (int64_t i = 0; i < m; ++i)
        tmp[i] = so[i] - sigma * softthres(so[i] / sigma - step * xg[i], delta * step);
      R nv = norm2(tmp.data(), m);
      R alpha = jl_max(R(0), R(1) - sl / nv);  // l2prox (:83)
      for (int64_t i = 0; i < m; ++i) y[b + i] = alpha * tmp[i];
    }
    for (int64_t i = b; i < e; ++i) y[i] -= (xk[i] + sj[i]);  // :116
  }
}

// ----------------------------------- ShiftedIndBallL0 / BInf (a16, a17) ----
// shiftedIndBallL0.jl:54-72, shiftedIndBallL0BInf.jl:73-95.
// sortperm!(p, y, rev=true, by=abs): indices ordered by |z| descending under
// isless (NaN largest), equal keys by ascending index.
template <class R>
void prox_indballl0(int64_t n, R* y, const R* xk, const R* sj, const R* q, int64_t r, bool binf,
                    R delta) {
  for (int64_t i = 0; i < n; ++i) y[i] = (xk[i] + sj[i]) + q[i];
  std::vector<int64_t> p((size_t)n);
  std::iota(p.begin(), p.end(), (int64_t)0);
  std::stable_sort(p.begin(), p.end(), [&](int64_t a, int64_t b) {
    return jl_isless((double)std::fabs(y[b]), (double)std::fabs(y[a]));
  });
  for (int64_t k = std::max<int64_t>(r, 0); k < n; ++k) y[p[(size_t)k]] = 0;
  if (!binf) {
    for (int64_t i = 0; i < n; ++i) y[i] -= (xk[i] + sj[i]);
  } else {
    for (int64_t i = 0; i < n; ++i)
      y[i] = jl_min(jl_max(y[i] - (xk[i] + sj[i]), -delta), delta);
  }
}

// ------------------------------------------------------- ψ(y) (a18-a20) ---
// h kinds: 0 NormL1, 1 NormL0, 2 RootNormLhalf, 3 IndBallL0 (param r).
// ProximalOperators 0.15 value functors restated; sums accumulated in extended
// precision (the reference's BLAS asum / pairwise mapreduce order is unspecified).
template <class R> double h_value(int kind, const R* v, int64_t m, R lambda, int64_t r) {
  const double inf = std::numeric_limits<double>::infinity();
  if (kind == 0) {
    long double s = 0;
    for (int64_t i = 0; i < m; ++i) s += std::fabs((long double)v[i]);
    return (double)(lambda * (R)s);
  } else if (kind == 1 || kind == 3) {
    int64_t c = 0;
    for (int64_t i = 0; i < m; ++i) c += (v[i] != R(0));
    if (kind == 1) return (double)(lambda * (R)c);
    return c <= r ? 0.0 : inf;
  } else {
    long double s = 0;  // rootNormLhalf.jl:27-29
    for (int64_t i = 0; i < m; ++i) s += (long double)std::sqrt(std::fabs(v[i]));
    return (double)(lambda * (R)s);
  }
}
// ShiftedProximalOperators.jl:51-54
template <class R>
double value_plain(int kind, int64_t n, R* xsy, const R* xk, const R* sj, const R* y, R lambda,
                   int64_t r) {
  for (int64_t i = 0; i < n; ++i) xsy[i] = (xk[i] + sj[i]) + y[i];
  return h_value(kind, xsy, n, lambda, r);
}
// shiftedNormL1Box.jl:70-82 (identical in L0Box :70-82 and LhalfBox :67-79)
template <class R>
double value_box(int kind, int64_t n, const R* xk, const R* sj, const R* y, Bound<R> l, Bound<R> u,
                 const Sel& sel, R lambda) {
  std::vector<R> xsy;
  if (sel.kind == 0) {
    xsy.resize((size_t)n);
    for (int64_t i = 0; i < n; ++i) xsy[i] = (xk[i] + sj[i]) + y[i];
  } else {
    xsy.resize((size_t)sel.nlist);
    for (int64_t j = 0; j < sel.nlist; ++j) {
      int64_t i = sel.list[j];
      xsy[j] = (xk[i] + sj[i]) + y[i];
    }
  }
  double val = h_value(kind, xsy.data(), (int64_t)xsy.size(), lambda, 0);
  R e = std::sqrt(std::numeric_limits<R>::epsilon());
  for (int64_t i = 0; i < n; ++i) {
    R w = sj[i] + y[i];
    if (!(l.at(i) - e <= w && w <= u.at(i) + e)) return std::numeric_limits<double>::infinity();
  }
  return val;
}
// shiftedNormL1B2.jl:32 : h(xk+sj+y) + IndBallL2(Δ)(sj+y).  IndBallL2 of
// ProximalOperators 0.15: 0 iff norm(x) <= r or isapprox(norm, r; atol=eps, rtol=√eps).
template <class R>
double value_l1b2(int64_t n, const R* xk, const R* sj, const R* y, R lambda, R delta) {
  std::vector<R> v((size_t)n), w((size_t)n);
  for (int64_t i = 0; i < n; ++i) { v[i] = (xk[i] + sj[i]) + y[i]; w[i] = sj[i] + y[i]; }
  double hv = h_value(0, v.data(), n, lambda, 0);
  R nw = norm2(w.data(), n);
  const R eps = std::numeric_limits<R>::epsilon();
  bool inside = (nw <= delta) ||
                (std::isfinite(nw) && std::isfinite(delta) &&
                 std::fabs(nw - delta) <= jl_max(eps, std::sqrt(eps) * jl_max(std::fabs(nw), std::fabs(delta))));
  return inside ? hv : std::numeric_limits<double>::infinity();
}
// shiftedIndBallL0BInf.jl:44-49 and shiftedGroupNormL2Binf.jl:34-39:
// w = sj + y; IndBallLinf(1.1Δ)(w) (strict: Inf iff some |w_i| > 1.1Δ, with
// 1.1 a Float64 literal); v = w + xk.  kind 3 -> IndBallL0, kind 4 -> GroupNormL2.
template <class R>
double value_binf(int kind, int64_t n, const R* xk, const R* sj, const R* y, R delta, int64_t r,
                  int64_t ngroups, const int64_t* offs, const R* lambda_g) {
  const double inf = std::numeric_limits<double>::infinity();
  double rad = 1.1 * (double)delta;  // Float64 radius even for Float32 data (1.1 is a Float64 literal)
  double ind = 0;
  std::vector<R> v((size_t)n);
  for (int64_t i = 0; i < n; ++i) {
    R w = sj[i] + y[i];
    if ((double)w < -rad || (double)w > rad) ind = inf;
    v[i] = w + xk[i];
  }
  double hv;
  if (kind == 3) {
    hv = h_value<R>(3, v.data(), n, R(0), r);
  } else {
    R s = 0;  // groupNormL2.jl:33-39
    for (int64_t g = 0; g < ngroups; ++g)
      s += lambda_g[g] * norm2(v.data() + offs[g], offs[g + 1] - offs[g]);
    hv = (double)s;
  }
  return hv + ind;
}
// generic ψ(y) for ShiftedGroupNormL2: ShiftedProximalOperators.jl:51-54 + groupNormL2.jl:33-39
template <class R>
double value_groupl2(int64_t n, const R* xk, const R* sj, const R* y, int64_t ngroups,
                     const int64_t* offs, const R* lambda_g) {
  std::vector<R> v((size_t)n);
  for (int64_t i = 0; i < n; ++i) v[i] = (xk[i] + sj[i]) + y[i];
  R s = 0;
  for (int64_t g = 0; g < ngroups; ++g)
    s += lambda_g[g] * norm2(v.data() + offs[g], offs[g + 1] - offs[g]);
  return (double)s;
}

}  // namespace

// ============================================================ C exports ====
#define ORC_API extern "C" __attribute__((visibility("default")))

#define ORC_DEFINE(SUF, R)                                                                         \
  ORC_API void orc_prox_l1_##SUF(int64_t n, R* y, const R* xk, const R* sj, const R* q,           \
                                 double lambda, double sigma) {                                   \
    prox_l1<R>(n, y, xk, sj, q, (R)lambda, (R)sigma);                                             \
  }                                                                                                \
  ORC_API int64_t orc_iprox_l1_##SUF(int64_t n, R* y, const R* xk, const R* sj, const R* g,       \
                                     const R* d, double lambda) {                                 \
    return iprox_l1<R>(n, y, xk, sj, g, d, (R)lambda);                                            \
  }                                                                                                \
  ORC_API void orc_prox_l0_##SUF(int64_t n, R* y, const R* xk, const R* sj, const R* q,           \
                                 double lambda, double sigma) {                                   \
    prox_l0<R>(n, y, xk, sj, q, (R)lambda, (R)sigma);                                             \
  }                                                                                                \
  ORC_API int64_t orc_iprox_l0_##SUF(int64_t n, R* y, const R* xk, const R* sj, const R* g,       \
                                     const R* d, double lambda) {                                 \
    return iprox_l0<R>(n, y, xk, sj, g, d, (R)lambda);                                            \
  }                                                                                                \
  ORC_API void orc_prox_lhalf_##SUF(int64_t n, R* y, R* sol, const R* xk, const R* sj,            \
                                    const R* q, double lambda, double sigma) {                    \
    prox_lhalf<R>(n, y, sol, xk, sj, q, (R)lambda, (R)sigma);                                     \
  }                                                                                                \
  ORC_API double orc_prox_rootlhalf_unshifted_##SUF(int64_t n, R* y, const R* x, double lambda,   \
                                                    double gamma) {                               \
    return prox_rootlhalf_unshifted<R>(n, y, x, (R)lambda, (R)gamma);                             \
  }                                                                                                \
  ORC_API void orc_prox_box_##SUF(int32_t op, int64_t n, R* y, R* sol, const R* xk, const R* sj,  \
                                  const R* q, const R* lvec, double lval, const R* uvec,          \
                                  double uval, int32_t sel_kind, const int64_t* sel_list,         \
                                  int64_t nsel, double lambda, double sigma) {                    \
    Sel sel(sel_kind, sel_list, nsel, n);                                                         \
    Bound<R> l{lvec, (R)lval}, u{uvec, (R)uval};                                                  \
    if (op == 0) prox_l1box<R>(n, y, xk, sj, q, l, u, sel, (R)lambda, (R)sigma);                  \
    else if (op == 1) prox_l0box<R>(n, y, xk, sj, q, l, u, sel, (R)lambda, (R)sigma);             \
    else prox_lhalfbox<R>(n, y, sol, xk, sj, q, l, u, sel, (R)lambda, (R)sigma);                  \
  }                                                                                                \
  ORC_API void orc_iprox_box_##SUF(int32_t op, int64_t n, R* y, const R* xk, const R* sj,         \
                                   const R* g, const R* d, const R* lvec, double lval,            \
                                   const R* uvec, double uval, int32_t sel_kind,                  \
                                   const int64_t* sel_list, int64_t nsel, double lambda) {        \
    Sel sel(sel_kind, sel_list, nsel, n);                                                         \
    Bound<R> l{lvec, (R)lval}, u{uvec, (R)uval};                                                  \
    if (op == 0) iprox_l1box<R>(n, y, xk, sj, g, d, l, u, sel, (R)lambda);                        \
    else iprox_l0box<R>(n, y, xk, sj, g, d, l, u, sel, (R)lambda);                                \
  }                                                                                                \
  ORC_API int32_t orc_prox_l1b2_##SUF(int64_t n, R* y, const R* xk, const R* sj, const R* q,      \
                                      double lambda, double sigma, double delta,                  \
                                      double chi_lambda) {                                        \
    return prox_l1b2<R>(n, y, xk, sj, q, (R)lambda, (R)sigma, (R)delta, (R)chi_lambda);           \
  }                                                                                                \
  ORC_API void orc_prox_groupl2_##SUF(int64_t n, R* y, R* sol, const R* xk, const R* sj,          \
                                      const R* q, int64_t ng, const int64_t* offs,                \
                                      const R* lambda, double sigma) {                            \
    prox_groupl2<R>(n, y, sol, xk, sj, q, ng, offs, lambda, (R)sigma);                            \
  }                                                                                                \
  ORC_API double orc_prox_groupl2_unshifted_##SUF(int64_t n, R* y, const R* x, int64_t ng,        \
                                                  const int64_t* offs, const R* lambda,           \
                                                  double gamma) {                                 \
    return prox_groupl2_unshifted<R>(n, y, x, ng, offs, lambda, (R)gamma);                        \
  }                                                                                                \
  ORC_API void orc_prox_groupl2binf_##SUF(int64_t n, R* y, R* sol, const R* xk, const R* sj,      \
                                          const R* q, int64_t ng, const int64_t* offs,            \
                                          const R* lambda, double sigma, double delta) {          \
    prox_groupl2binf<R>(n, y, sol, xk, sj, q, ng, offs, lambda, (R)sigma, (R)delta);              \
  }                                                                                                \
  ORC_API void orc_prox_groupl2binf_dbg_##SUF(int64_t n, R* y, R* sol, const R* xk, const R* sj,  \
                                              const R* q, int64_t ng, const int64_t* offs,        \
                                              const R* lambda, double sigma, double delta,        \
                                              double* nroot, int8_t* zero_g, int64_t ulp_shift) { \
    prox_groupl2binf<R>(n, y, sol, xk, sj, q, ng, offs, lambda, (R)sigma, (R)delta, nroot,        \
                        zero_g, ulp_shift);                                                        \
  }                                                                                                \
  ORC_API void orc_prox_lhalfbox_dbg_##SUF(int64_t n, R* y, R* sol, const R* xk, const R* sj,     \
                                           const R* q, const R* lvec, double lval, const R* uvec, \
                                           double uval, int32_t sel_kind, const int64_t* sel_list,\
                                           int64_t nsel, double lambda, double sigma,             \
                                           int8_t* pick, double* obj) {                           \
    Sel sel(sel_kind, sel_list, nsel, n);                                                         \
    Bound<R> l{lvec, (R)lval}, u{uvec, (R)uval};                                                  \
    prox_lhalfbox<R>(n, y, sol, xk, sj, q, l, u, sel, (R)lambda, (R)sigma, pick, obj);            \
  }                                                                                                \
  ORC_API void orc_prox_indballl0_##SUF(int64_t n, R* y, const R* xk, const R* sj, const R* q,    \
                                        int64_t r, int32_t binf, double delta) {                  \
    prox_indballl0<R>(n, y, xk, sj, q, r, binf != 0, (R)delta);                                   \
  }                                                                                                \
  ORC_API double orc_value_plain_##SUF(int32_t kind, int64_t n, R* xsy, const R* xk,              \
                                       const R* sj, const R* y, double lambda, int64_t r) {       \
    return value_plain<R>(kind, n, xsy, xk, sj, y, (R)lambda, r);                                 \
  }                                                                                                \
  ORC_API double orc_value_box_##SUF(int32_t kind, int64_t n, const R* xk, const R* sj,           \
                                     const R* y, const R* lvec, double lval, const R* uvec,       \
                                     double uval, int32_t sel_kind, const int64_t* sel_list,      \
                                     int64_t nsel, double lambda) {                               \
    Sel sel(sel_kind, sel_list, nsel, n);                                                         \
    Bound<R> l{lvec, (R)lval}, u{uvec, (R)uval};                                                  \
    return value_box<R>(kind, n, xk, sj, y, l, u, sel, (R)lambda);                                \
  }                                                                                                \
  ORC_API double orc_value_l1b2_##SUF(int64_t n, const R* xk, const R* sj, const R* y,            \
                                      double lambda, double delta) {                              \
    return value_l1b2<R>(n, xk, sj, y, (R)lambda, (R)delta);                                      \
  }                                                                                                \
  ORC_API double orc_value_binf_##SUF(int32_t kind, int64_t n, const R* xk, const R* sj,          \
                                      const R* y, double delta, int64_t r, int64_t ng,            \
                                      const int64_t* offs, const R* lambda_g) {                   \
    return value_binf<R>(kind, n, xk, sj, y, (R)delta, r, ng, offs, lambda_g);                    \
  }                                                                                                \
  ORC_API double orc_value_groupl2_##SUF(int64_t n, const R* xk, const R* sj, const R* y,         \
                                         int64_t ng, const int64_t* offs, const R* lambda_g) {    \
    return value_groupl2<R>(n, xk, sj, y, ng, offs, lambda_g);                                    \
  }                                                                                                \
  ORC_API double orc_prox_zero_##SUF(double q, double l, double u) {                              \
    return (double)prox_zero<R>((R)q, (R)l, (R)u);                                                \
  }                                                                                                \
  ORC_API double orc_iprox_zero_##SUF(double d, double g, double l, double u) {                   \
    return (double)iprox_zero<R>((R)d, (R)g, (R)l, (R)u);                                         \
  }

ORC_DEFINE(f64, double)
ORC_DEFINE(f32, float)

// Synthetic inputs shared bit-for-bit by host and device:
// u(i,k) = (splitmix64(seed ^ (k<<40) + i) >> 11) * 2^-53   (f32: >> 40, * 2^-24)
static inline uint64_t splitmix64(uint64_t x) {
  x += 0x9E3779B97F4A7C15ull;
  x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
  x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
  return x ^ (x >> 31);
}
ORC_API void orc_fill_uniform_f64(double* out, int64_t n, int64_t i0, uint64_t seed, uint64_t stream,
                                  double scale, double shift) {
  for (int64_t i = 0; i < n; ++i) {
    uint64_t h = splitmix64((seed ^ (stream << 40)) + (uint64_t)(i0 + i));
    out[i] = scale * ((double)(h >> 11) * 0x1.0p-53) + shift;
  }
}
ORC_API void orc_fill_uniform_f32(float* out, int64_t n, int64_t i0, uint64_t seed, uint64_t stream,
                                  float scale, float shift) {
  for (int64_t i = 0; i < n; ++i) {
    uint64_t h = splitmix64((seed ^ (stream << 40)) + (uint64_t)(i0 + i));
    out[i] = scale * ((float)(h >> 40) * 0x1.0p-24f) + shift;
  }
}
ORC_API int32_t orc_version(void) { return 1; }
