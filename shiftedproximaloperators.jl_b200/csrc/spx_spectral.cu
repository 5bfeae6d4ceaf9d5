// spx_spectral.cu -- the thresholding stage of the spectral operators, given an SVD.
//
// ShiftedRank / ShiftedNuclearnorm / ShiftedCappedl1 prox! (shiftedRank.jl:68-84, shiftedNuclearnorm.jl:68-81,
// shiftedCappedl1.jl:68-86) are
//     sol = q + xk + sj;  A = reshape(sol);  U, S, Vt = svd(A)          <- LAPACK in the reference: out of scope
//     S'  = threshold(S);  U[:, i] *= S'_i                              <- this file
//     A   = U * Vt                                                      <- BLAS `mul!`: a library GEMM, out of scope
//     y   = reshape(A) - (xk + sj)                                      <- this file
// The SVD and the GEMM stay library calls on the caller's side (cuSOLVER / cuBLAS); the three elementwise stages
// around them run here, bit-identical to the reference loops: one multiplication per entry of U, the reference's
// comparison operators for the three thresholds.
#include "spx_elementwise.cuh"

namespace spx {

// sol = (q + xk) + sj   (`ψ.sol .= q .+ ψ.xk .+ ψ.sj`, shiftedRank.jl:69)
template <class R> struct SpectralSol {
  using Real = R;
  static constexpr int NIN = 3, UNROLL = 4;
  static constexpr bool OUT = true, ACC = false;
  const R* in[NIN];  // xk, sj, q
  R fill[NIN];
  R* y;
  __device__ __forceinline__ R apply(const R (&x)[NIN], long long, Partial&) const { return (x[2] + x[0]) + x[1]; }
};
// y = A - (xk + sj)   (shiftedRank.jl:82)
template <class R> struct SpectralFinish {
  using Real = R;
  static constexpr int NIN = 3, UNROLL = 4;
  static constexpr bool OUT = true, ACC = false;
  const R* in[NIN];  // A, xk, sj
  R fill[NIN];
  R* y;
  __device__ __forceinline__ R apply(const R (&x)[NIN], long long, Partial&) const { return x[0] - (x[1] + x[2]); }
};

// the factor column i of U is multiplied by, and the value left in S[i]
//   kind 0 Rank       (shiftedRank.jl:72-80):        S_i <= sqrt(2λσ) ? column zeroed : column * S_i;   S untouched
//   kind 1 Nuclearnorm (shiftedNuclearnorm.jl:72-77): S_i = max(0, S_i - λσ);  column * S_i
//   kind 2 Cappedl1   (shiftedCappedl1.jl:71-82):    x1 = max(θ, S_i), x2 = min(θ, max(0, S_i - λσ));
//          S_i = ((x1-S_i)²/2 + λσθ < (x2-S_i)²/2 + λσ x2) ? x1 : x2;  column * S_i
template <class R> __device__ __forceinline__ R spectral_factor(int kind, R s, R lambda, R sigma, R theta, bool& zero_col) {
  zero_col = false;
  if (kind == 0) {
    const R c = sqrt(R(2) * lambda * sigma);
    zero_col = s <= c;
    return s;
  }
  if (kind == 1) return jl_max(R(0), s - lambda * sigma);
  const R x1 = jl_max(theta, s);
  const R x2 = jl_min(theta, jl_max(R(0), s - lambda * sigma));
  const R d1 = x1 - s, d2 = x2 - s;
  const R v1 = (d1 * d1) / R(2) + lambda * sigma * theta;
  const R v2 = (d2 * d2) / R(2) + lambda * sigma * x2;
  return v1 < v2 ? x1 : x2;
}

// U is m x k column-major with leading dimension ldu; one CTA row of 256 threads per 256 rows, grid.y over columns
template <class R>
__global__ void __launch_bounds__(256) spectral_scale_kernel(int kind, long long m, long long k, R* __restrict__ U,
                                                             long long ldu, const R* __restrict__ S, R lambda, R sigma,
                                                             R theta) {
  for (long long i = blockIdx.y; i < k; i += gridDim.y) {
    bool zero_col;
    const R f = spectral_factor<R>(kind, S[i], lambda, sigma, theta, zero_col);
    R* col = U + i * ldu;
    for (long long j = (long long)blockIdx.x * 256 + threadIdx.x; j < m; j += (long long)gridDim.x * 256)
      col[j] = zero_col ? R(0) : col[j] * f;  // `U[:, i] .= 0` / `U[j, i] * S[i]`
  }
}
// S is rewritten after every column has read it (Nuclearnorm and Cappedl1 store the thresholded values)
template <class R>
__global__ void __launch_bounds__(256) spectral_store_s_kernel(int kind, long long k, R* __restrict__ S, R lambda, R sigma,
                                                               R theta) {
  const long long i = (long long)blockIdx.x * 256 + threadIdx.x;
  if (i < k && kind != 0) {
    bool z;
    S[i] = spectral_factor<R>(kind, S[i], lambda, sigma, theta, z);
  }
}

template <class Op, class R> static int32_t run3(spx_ctx* ctx, int64_t n, R* out, const R* a, const R* b, const R* c) {
  SPX_REQUIRE(ctx != nullptr, "null context");
  SPX_REQUIRE(n >= 0, "n < 0");
  SPX_REQUIRE(n == 0 || (out && a && b && c), "null device vector");
  DeviceGuard g(ctx->device);
  Op op;
  op.in[0] = a; op.in[1] = b; op.in[2] = c;
  op.fill[0] = op.fill[1] = op.fill[2] = R(0);
  op.y = out;
  return ew_launch(ctx, ctx->stream, op, n, 0, ctx->d_partials, nullptr);
}

template <class R>
static int32_t spectral_threshold(spx_ctx* ctx, int32_t kind, int64_t m, int64_t k, R* U, int64_t ldu, R* S, double lambda,
                                  double sigma, double theta) {
  SPX_REQUIRE(ctx != nullptr, "null context");
  SPX_REQUIRE(kind >= 0 && kind <= 2, "kind: 0 Rank, 1 Nuclearnorm, 2 Cappedl1");
  SPX_REQUIRE(m >= 0 && k >= 0 && ldu >= m, "bad matrix shape");
  SPX_REQUIRE((m == 0 || k == 0) || (U && S), "null device pointer");
  if (m == 0 || k == 0) return SPX_OK;
  DeviceGuard g(ctx->device);
  const unsigned gx = (unsigned)std::min<int64_t>((m + 255) / 256, 4096);
  const unsigned gy = (unsigned)std::min<int64_t>(k, 65535);
  spectral_scale_kernel<R><<<dim3(gx, gy), 256, 0, ctx->stream>>>(kind, m, k, U, ldu, S, (R)lambda, (R)sigma, (R)theta);
  spectral_store_s_kernel<R><<<(unsigned)((k + 255) / 256), 256, 0, ctx->stream>>>(kind, k, S, (R)lambda, (R)sigma, (R)theta);
  ctx->launches += 2;
  SPX_CUDA(cudaGetLastError());
  return SPX_OK;
}

}  // namespace spx

using namespace spx;

#define SPX_DEFINE_SPECTRAL(SUF, R)                                                                                    \
  extern "C" int32_t spx_spectral_sol_##SUF(spx_ctx* ctx, int64_t n, R* a_out, const R* xk, const R* sj, const R* q) { \
    return run3<SpectralSol<R>, R>(ctx, n, a_out, xk, sj, q);                                                          \
  }                                                                                                                    \
  extern "C" int32_t spx_spectral_threshold_##SUF(spx_ctx* ctx, int32_t kind, int64_t m, int64_t k, R* u, int64_t ldu, \
                                                  R* s, double lambda, double sigma, double theta) {                  \
    return spectral_threshold<R>(ctx, kind, m, k, u, ldu, s, lambda, sigma, theta);                                    \
  }                                                                                                                    \
  extern "C" int32_t spx_spectral_finish_##SUF(spx_ctx* ctx, int64_t n, R* y, const R* a, const R* xk, const R* sj) { \
    return run3<SpectralFinish<R>, R>(ctx, n, y, a, xk, sj);                                                           \
  }

SPX_DEFINE_SPECTRAL(f64, double)
SPX_DEFINE_SPECTRAL(f32, float)
