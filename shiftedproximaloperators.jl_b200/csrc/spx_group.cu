// spx_group.cu -- group-norm operators: segmented warp-shuffle reductions.
//
// One warp owns one group at a time (grid-stride over groups, so neighbouring
// warps stream neighbouring groups: every 128-byte line is consumed while it is
// in flight).  Groups of up to 32*EPL elements live in registers for the whole
// prox (one HBM read per operand, one write); longer groups stash `sol` in the
// output vector and re-read it through L1/L2.  Sums of squares are accumulated
// in Float64 whatever R is; the butterfly order is fixed, so results are
// deterministic.
#include "spx_common.cuh"

namespace spx {

constexpr int kGroupThreads = 256;
constexpr int kEPL = 4;  // elements per lane kept in registers (groups <= 128 elements)

template <class R> __device__ __forceinline__ R ldv(const R* p) {
  Pack<R, 1> t;
  ld_stream(p, t);
  return t.v[0];
}
template <class R> __device__ __forceinline__ void stv(R* p, R v) {
  Pack<R, 1> t;
  t.v[0] = v;
  st_stream(p, t);
}

template <class R> __device__ __forceinline__ R softthres(R x, R a) {
  return jl_sign(x) * jl_max(R(0), jl_abs(x) - a);
}

// ------------------------------------------------------- ShiftedGroupNormL2 --
// shiftedGroupNormL2.jl:52-79
template <class R, bool PSI>
__global__ void __launch_bounds__(kGroupThreads)
    group_l2_kernel(R* y, const R* xk, const R* sj, const R* q, long long ngroups,
                    const long long* __restrict__ offs, const R* __restrict__ lambda_g, R sigma,
                    Partial* __restrict__ partials) {
  const int lane = threadIdx.x & 31;
  const long long warp = ((long long)blockIdx.x * kGroupThreads + threadIdx.x) >> 5;
  const long long nwarps = ((long long)gridDim.x * kGroupThreads) >> 5;
  double psi = 0.0;
  for (long long g = warp; g < ngroups; g += nwarps) {
    const long long b = offs[g], e = offs[g + 1], m = e - b;
    const R lam = lambda_g[g];
    double vv = 0.0;
    if (m <= 32 * kEPL) {
      R sol[kEPL], xs[kEPL];
      double ss = 0.0;
#pragma unroll
      for (int k = 0; k < kEPL; ++k) {
        const long long i = b + k * 32 + lane;
        sol[k] = R(0);
        xs[k] = R(0);
        if (i < e) {
          const R xi = ldv(xk + i), si = ldv(sj + i), qi = ldv(q + i);
          sol[k] = (qi + xi) + si;  // :65
          xs[k] = xi + si;
          ss += (double)sol[k] * (double)sol[k];
        }
      }
      ss = warp_sum(ss);
      const R snorm = (R)sqrt(ss);
      const R alpha = jl_max(R(1) - sigma * lam / snorm, R(0));
#pragma unroll
      for (int k = 0; k < kEPL; ++k) {
        const long long i = b + k * 32 + lane;
        if (i < e) {
          const R o = (snorm == R(0) ? R(0) : alpha * sol[k]) - xs[k];  // :70-77
          stv(y + i, o);
          if (PSI) {
            const double v = (double)(xs[k] + o);
            vv += v * v;
          }
        }
      }
    } else {
      double ss = 0.0;
      for (long long i = b + lane; i < e; i += 32) {
        const R s = (q[i] + xk[i]) + sj[i];
        y[i] = s;  // stash sol (each lane re-reads only what it wrote)
        ss += (double)s * (double)s;
      }
      ss = warp_sum(ss);
      const R snorm = (R)sqrt(ss);
      const R alpha = jl_max(R(1) - sigma * lam / snorm, R(0));
      for (long long i = b + lane; i < e; i += 32) {
        const R xsi = xk[i] + sj[i];
        const R o = (snorm == R(0) ? R(0) : alpha * y[i]) - xsi;
        y[i] = o;
        if (PSI) {
          const double v = (double)(xsi + o);
          vv += v * v;
        }
      }
    }
    if (PSI) {
      vv = warp_sum(vv);
      if (lane == 0) psi += (double)(lam * (R)sqrt(vv));  // λ_g ‖v_g‖  groupNormL2.jl:36
    }
  }
  if (PSI) {
    Partial p;
    p.s = psi;
    p.s2 = 0.0;
    p.bad = -1;
    p = block_fold<kGroupThreads>(p);
    if (threadIdx.x == 0) partials[blockIdx.x] = p;
  }
}

// --------------------------------------------------- ShiftedGroupNormL2Binf --
// shiftedGroupNormL2Binf.jl:67-119.  Group data access: registers (REG) or the
// stashed `sol` in y plus xk from memory.
template <class R> struct GroupView {
  bool reg;
  R sol[kEPL], xkr[kEPL];
  const R* ysol;  // stashed sol (y)
  const R* xk;
  long long b, e;
  int lane;
  // Σ f(sol_i, xk_i)^2 over the group, Float64 accumulate, all lanes get the sum
  template <class F> __device__ __forceinline__ double sumsq(F f) const {
    double ss = 0.0;
    if (reg) {
#pragma unroll
      for (int k = 0; k < kEPL; ++k) {
        const long long i = b + k * 32 + lane;
        if (i < e) {
          const double t = (double)f(sol[k], xkr[k]);
          ss += t * t;
        }
      }
    } else {
      for (long long i = b + lane; i < e; i += 32) {
        const double t = (double)f(ysol[i], xk[i]);
        ss += t * t;
      }
    }
    return warp_sum(ss);
  }
};

template <class R> __device__ __forceinline__ bool adjacent_or_crossed(R a, R m, R b) { return !((a < m) && (m < b)); }

template <class R>
__global__ void __launch_bounds__(kGroupThreads)
    group_l2binf_kernel(R* y, const R* xk, const R* sj, const R* q, long long ngroups,
                        const long long* __restrict__ offs, const R* __restrict__ lambda_g, R sigma, R delta) {
  const int lane = threadIdx.x & 31;
  const long long warp = ((long long)blockIdx.x * kGroupThreads + threadIdx.x) >> 5;
  const long long nwarps = ((long long)gridDim.x * kGroupThreads) >> 5;
  const R eps = Eps<R>::value;
  for (long long g = warp; g < ngroups; g += nwarps) {
    GroupView<R> gv;
    gv.b = offs[g];
    gv.e = offs[g + 1];
    gv.lane = lane;
    gv.ysol = y;
    gv.xk = xk;
    gv.reg = (gv.e - gv.b) <= 32 * kEPL;
    const R lam = lambda_g[g];
    R xs[kEPL];
    if (gv.reg) {
#pragma unroll
      for (int k = 0; k < kEPL; ++k) {
        const long long i = gv.b + k * 32 + lane;
        gv.sol[k] = R(0);
        gv.xkr[k] = R(0);
        xs[k] = R(0);
        if (i < gv.e) {
          const R xi = ldv(xk + i), si = ldv(sj + i), qi = ldv(q + i);
          gv.sol[k] = (qi + xi) + si;  // :80
          gv.xkr[k] = xi;
          xs[k] = xi + si;
        }
      }
    } else {
      for (long long i = gv.b + lane; i < gv.e; i += 32) y[i] = (q[i] + xk[i]) + sj[i];
    }
    const R sl = lam * sigma;  // σλ
    auto cstep = [&](R nn) -> R { return nn / (sigma * (nn - sl)); };
    auto froot = [&](R nn) -> R {  // :87-93
      const R c = cstep(nn);
      const R dc = delta * c;
      const double ss =
          gv.sumsq([&](R so, R xg) -> R { return sigma * softthres(so / sigma - c * xg, dc) - so; });
      return nn - (R)sqrt(ss);
    };
    const R lmin = sl * (R(1) + eps);
    const R fl = froot(lmin);
    const R ansatz = lmin + R(1);
    R step = ansatz / (sigma * (ansatz - sl));
    const R dstep = delta * step;
    const R zlmax = (R)sqrt(gv.sumsq([&](R so, R xg) -> R { return softthres(so / sigma - step * xg, dstep); }));
    const R nsol = (R)sqrt(gv.sumsq([&](R so, R) -> R { return so; }));
    const R nxk = (R)sqrt(gv.sumsq([&](R, R xg) -> R { return xg; }));
    const R lmax = nsol + sigma * (zlmax + R(1) * lam * nxk);  // |(ϵ-1)/ϵ + 1| = 1 for ϵ = 1  (:100)
    const R fm = froot(lmax);
    bool zero_out = false;
    R nroot = R(0);
    if (fl * fm > R(0)) {
      zero_out = true;
    } else {
      // fzero(froot, lmin, lmax): Roots' bisection ends on two adjacent floats
      // around the sign change.  Same end state, fewer evaluations: Illinois
      // regula falsi steps while they at least halve the bracket, bisection
      // otherwise, until lo and hi are adjacent.
      R a = lmin, fa = fl, bb = lmax, fb = fm;
      if (fa == R(0)) {
        nroot = a;
      } else if (fb == R(0)) {
        nroot = bb;
      } else {
        R ga = fa, gb = fb;  // Illinois-damped copies
        int side = 0;
        bool force_bisect = false;
        for (int it = 0; it < 200; ++it) {
          const R mid = a + (bb - a) / R(2);
          if (adjacent_or_crossed(a, mid, bb)) break;
          R x = mid;
          if (!force_bisect) {
            const R xs_ = (a * gb - bb * ga) / (gb - ga);
            if ((a < xs_) && (xs_ < bb)) x = xs_;
          }
          const R width = bb - a;
          const R fx = froot(x);
          if (fx == R(0)) {
            a = bb = x;
            fa = fb = R(0);
            break;
          }
          if ((fx < R(0)) == (fa < R(0))) {
            a = x; fa = fx; ga = fx;
            if (side == -1) gb = gb / R(2);
            side = -1;
          } else {
            bb = x; fb = fx; gb = fx;
            if (side == 1) ga = ga / R(2);
            side = 1;
          }
          force_bisect = !((bb - a) <= width / R(2));
          if (force_bisect) { ga = fa; gb = fb; side = 0; }
        }
        nroot = (jl_abs(fa) <= jl_abs(fb)) ? a : bb;
      }
      step = cstep(nroot);
      if (jl_abs(nroot - sl) == R(0)) zero_out = true;  // `abs(n - σλ) ≈ 0`  (:107)
    }
    // y_g = l2prox(sol - σ softthres(sol/σ - step xk, Δ step), σλ) - (xk + sj)   (:109-116)
    const R dstep2 = delta * step;
    R alpha = R(0);
    if (!zero_out) {
      const double ss =
          gv.sumsq([&](R so, R xg) -> R { return so - sigma * softthres(so / sigma - step * xg, dstep2); });
      alpha = jl_max(R(0), R(1) - sl / (R)sqrt(ss));
    }
    if (gv.reg) {
#pragma unroll
      for (int k = 0; k < kEPL; ++k) {
        const long long i = gv.b + k * 32 + lane;
        if (i < gv.e) {
          R o = R(0);
          if (!zero_out) o = alpha * (gv.sol[k] - sigma * softthres(gv.sol[k] / sigma - step * gv.xkr[k], dstep2));
          o = o - xs[k];
          stv(y + i, o);
        }
      }
    } else {
      for (long long i = gv.b + lane; i < gv.e; i += 32) {
        R o = R(0);
        const R so = y[i], xg = xk[i];
        if (!zero_out) o = alpha * (so - sigma * softthres(so / sigma - step * xg, dstep2));
        y[i] = o - (xg + sj[i]);
      }
    }
  }
}

// ----------------------------------------------------- group values (ψ(y)) --
// ShiftedGroupNormL2: v = (xk + sj) + y  (ShiftedProximalOperators.jl:51-54)
// ShiftedGroupNormL2Binf: w = sj + y, IndBallLinf(1.1Δ)(w), v = w + xk
// (shiftedGroupNormL2Binf.jl:34-39); value Σ_g λ_g ‖v_g‖  (groupNormL2.jl:33-39)
template <class R>
__global__ void __launch_bounds__(kGroupThreads)
    group_value_kernel(const R* __restrict__ xk, const R* __restrict__ sj, const R* __restrict__ y, long long ngroups,
                       const long long* __restrict__ offs, const R* __restrict__ lambda_g, bool binf, double rad,
                       Partial* __restrict__ partials) {
  const int lane = threadIdx.x & 31;
  const long long warp = ((long long)blockIdx.x * kGroupThreads + threadIdx.x) >> 5;
  const long long nwarps = ((long long)gridDim.x * kGroupThreads) >> 5;
  Partial p;
  p.s = 0.0;
  p.s2 = 0.0;
  p.bad = -1;
  for (long long g = warp; g < ngroups; g += nwarps) {
    const long long b = offs[g], e = offs[g + 1];
    double ss = 0.0;
    for (long long i = b + lane; i < e; i += 32) {
      R v;
      if (binf) {
        const R w = sj[i] + y[i];
        if ((double)w < -rad || (double)w > rad) p.bad = 1;
        v = w + xk[i];
      } else {
        v = (xk[i] + sj[i]) + y[i];
      }
      ss += (double)v * (double)v;
    }
    ss = warp_sum(ss);
    if (lane == 0) p.s += (double)(lambda_g[g] * (R)sqrt(ss));
  }
  p = block_fold<kGroupThreads>(p);
  if (threadIdx.x == 0) partials[blockIdx.x] = p;
}

static int group_grid(spx_ctx* ctx, int64_t ngroups, const void* kernel) {
  int per_sm = 1;
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, kGroupThreads, 0) != cudaSuccess || per_sm < 1)
    per_sm = 1;
  long long want = (ngroups + (kGroupThreads / 32) - 1) / (kGroupThreads / 32);
  long long cap = (long long)ctx->sm_count * per_sm;
  if (cap > kMaxPartials) cap = kMaxPartials;
  if (want < 1) want = 1;
  return (int)(want < cap ? want : cap);
}

template <class R>
int32_t value_group_binf(spx_ctx* ctx, int64_t n, const R* xk, const R* sj, const R* y, bool binf, double delta,
                         int64_t ngroups, const int64_t* offs, const R* lambda_g, double* out) {
  SPX_REQUIRE(ctx && out, "null argument");
  SPX_REQUIRE(n >= 0 && ngroups >= 0, "negative size");
  SPX_REQUIRE(ngroups == 0 || (offs && lambda_g && xk && sj && y), "null device vector");
  DeviceGuard g(ctx->device);
  if (ngroups == 0) {
    *out = 0.0;
    return SPX_OK;
  }
  const int grid = group_grid(ctx, ngroups, (const void*)group_value_kernel<R>);
  group_value_kernel<R><<<grid, kGroupThreads, 0, ctx->stream>>>(xk, sj, y, ngroups, (const long long*)offs, lambda_g,
                                                                binf, 1.1 * (double)(R)delta, ctx->d_partials);
  ctx->launches++;
  SPX_CUDA(cudaGetLastError());
  int32_t st = finalize_partials(ctx, grid, 1, false);
  if (st != SPX_OK) return st;
  // the reference accumulates sum_c in R; one rounding to R here
  *out = ctx->h_result[0].bad > 0 ? std::numeric_limits<double>::infinity() : (double)(R)ctx->h_result[0].s;
  return SPX_OK;
}
template int32_t value_group_binf<double>(spx_ctx*, int64_t, const double*, const double*, const double*, bool, double,
                                          int64_t, const int64_t*, const double*, double*);
template int32_t value_group_binf<float>(spx_ctx*, int64_t, const float*, const float*, const float*, bool, double,
                                         int64_t, const int64_t*, const float*, double*);

template <class R>
static int32_t prox_group(spx_ctx* ctx, bool binf, int64_t n, R* y, const R* xk, const R* sj, const R* q,
                          int64_t ngroups, const int64_t* offs, const R* lambda_g, double sigma, double delta,
                          double* psi_out) {
  SPX_REQUIRE(ctx != nullptr, "null context");
  SPX_REQUIRE(n >= 0 && ngroups >= 0, "negative size");
  SPX_REQUIRE(ngroups == 0 || (y && xk && sj && q && offs && lambda_g), "null device vector");
  DeviceGuard g(ctx->device);
  if (ngroups > 0) {
    if (!binf) {
      if (psi_out) {
        const int grid = group_grid(ctx, ngroups, (const void*)group_l2_kernel<R, true>);
        group_l2_kernel<R, true><<<grid, kGroupThreads, 0, ctx->stream>>>(
            y, xk, sj, q, ngroups, (const long long*)offs, lambda_g, (R)sigma, ctx->d_partials);
        ctx->launches++;
        SPX_CUDA(cudaGetLastError());
        int32_t st = finalize_partials(ctx, grid, 1, false);
        if (st != SPX_OK) return st;
        *psi_out = (double)(R)ctx->h_result[0].s;
        return SPX_OK;
      }
      const int grid = group_grid(ctx, ngroups, (const void*)group_l2_kernel<R, false>);
      group_l2_kernel<R, false><<<grid, kGroupThreads, 0, ctx->stream>>>(
          y, xk, sj, q, ngroups, (const long long*)offs, lambda_g, (R)sigma, ctx->d_partials);
      ctx->launches++;
      SPX_CUDA(cudaGetLastError());
      return SPX_OK;
    }
    const int grid = group_grid(ctx, ngroups, (const void*)group_l2binf_kernel<R>);
    group_l2binf_kernel<R><<<grid, kGroupThreads, 0, ctx->stream>>>(
        y, xk, sj, q, ngroups, (const long long*)offs, lambda_g, (R)sigma, (R)delta);
    ctx->launches++;
    SPX_CUDA(cudaGetLastError());
  }
  if (psi_out) {
    if (!binf) {
      *psi_out = 0.0;
      return SPX_OK;
    }
    // the trust-region value needs sj + y and xk separately: one more streaming pass
    return value_group_binf<R>(ctx, n, xk, sj, y, true, delta, ngroups, offs, lambda_g, psi_out);
  }
  return SPX_OK;
}

}  // namespace spx

using namespace spx;

#define SPX_DEFINE_GROUP(SUF, R)                                                                                 \
  extern "C" int32_t spx_prox_groupl2_##SUF(spx_ctx* ctx, int64_t n, R* y, const R* xk, const R* sj, const R* q, \
                                            int64_t ngroups, const int64_t* offs, const R* lambda_g,             \
                                            double sigma, double* psi_out) {                                     \
    return prox_group<R>(ctx, false, n, y, xk, sj, q, ngroups, offs, lambda_g, sigma, 0.0, psi_out);             \
  }                                                                                                              \
  extern "C" int32_t spx_prox_groupl2binf_##SUF(spx_ctx* ctx, int64_t n, R* y, const R* xk, const R* sj,         \
                                                const R* q, int64_t ngroups, const int64_t* offs,                \
                                                const R* lambda_g, double sigma, double delta,                   \
                                                double* psi_out) {                                               \
    return prox_group<R>(ctx, true, n, y, xk, sj, q, ngroups, offs, lambda_g, sigma, delta, psi_out);            \
  }                                                                                                              \
  extern "C" int32_t spx_value_groupl2_##SUF(spx_ctx* ctx, int64_t n, const R* xk, const R* sj, const R* y,      \
                                             int64_t ngroups, const int64_t* offs, const R* lambda_g,            \
                                             double* out) {                                                      \
    return value_group_binf<R>(ctx, n, xk, sj, y, false, 0.0, ngroups, offs, lambda_g, out);                     \
  }

SPX_DEFINE_GROUP(f64, double)
SPX_DEFINE_GROUP(f32, float)
