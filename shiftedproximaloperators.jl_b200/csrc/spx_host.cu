// spx_host.cu -- Box prox!/iprox! with every vector in HOST memory (the
// end-to-end path bench.py times).  The vectors are cut into chunks; each chunk
// travels H2D -> fused kernel -> D2H on one of three streams, so the two copy
// engines and the SMs work on different chunks at the same time.  Device
// staging is a ring of three slots owned by the context.
#include <vector>

#include "spx_common.cuh"

namespace spx {

constexpr int kSlots = 3;

static int32_t ensure_pipe(spx_ctx* ctx, size_t bytes) {
  for (int i = 0; i < kSlots; ++i)
    if (!ctx->pipe_streams[i]) SPX_CUDA(cudaStreamCreateWithFlags(&ctx->pipe_streams[i], cudaStreamNonBlocking));
  if (ctx->pipe_bytes < bytes) {
    if (ctx->pipe_buf) {
      for (int i = 0; i < kSlots; ++i) SPX_CUDA(cudaStreamSynchronize(ctx->pipe_streams[i]));
      SPX_CUDA(cudaFree(ctx->pipe_buf));
      ctx->pipe_buf = nullptr;
      ctx->pipe_bytes = 0;
    }
    SPX_CUDA(cudaMalloc(&ctx->pipe_buf, bytes));
    ctx->pipe_bytes = bytes;
  }
  return SPX_OK;
}

template <class R>
static int32_t box_host(spx_ctx* ctx, int32_t op, int64_t n, R* y_h, const R* xk_h, const R* sj_h, const R* qg_h,
                        const R* d_h, const R* l_h, double l_val, const R* u_h, double u_val, double lambda,
                        double sigma, int64_t chunk, double* psi_out) {
  SPX_REQUIRE(ctx != nullptr, "null context");
  SPX_REQUIRE(n >= 0, "n < 0");
  SPX_REQUIRE(n == 0 || (y_h && xk_h && sj_h && qg_h), "null host vector");
  SPX_REQUIRE(op >= BOX_L1 && op <= BOX_LHALF, "unknown Box operator");
  SPX_REQUIRE(!(d_h != nullptr && op == BOX_LHALF), "RootNormLhalfBox has no iprox!");
  DeviceGuard guard(ctx->device);
  if (chunk <= 0) chunk = (int64_t)1 << 22;
  chunk = (chunk + 3) & ~(int64_t)3;  // keeps every chunk pointer 16-byte aligned
  if (chunk > n) chunk = ((n + 3) & ~(int64_t)3);
  if (n == 0) {
    if (psi_out) *psi_out = 0.0;
    return SPX_OK;
  }
  const bool inverse = d_h != nullptr;
  const int nvec_in = 3 + (inverse ? 1 : 0) + (l_h ? 1 : 0) + (u_h ? 1 : 0);
  const size_t vec_bytes = (size_t)chunk * sizeof(R);
  const size_t slot_bytes = vec_bytes * (size_t)(nvec_in + 1);
  const int64_t nchunks = (n + chunk - 1) / chunk;
  const size_t res_bytes = sizeof(Partial) * (size_t)nchunks;
  int32_t st = ensure_pipe(ctx, slot_bytes * kSlots + res_bytes + 256);
  if (st != SPX_OK) return st;
  char* base = (char*)ctx->pipe_buf;
  Partial* d_res = (Partial*)(base + slot_bytes * kSlots);
  DevSel sel;
  make_sel(nullptr, n, &sel);
  for (int64_t c = 0; c < nchunks; ++c) {
    const int slot = (int)(c % kSlots);
    cudaStream_t s = ctx->pipe_streams[slot];
    const int64_t i0 = c * chunk;
    const int64_t m = (n - i0) < chunk ? (n - i0) : chunk;
    const size_t bytes = (size_t)m * sizeof(R);
    R* buf = (R*)(base + slot_bytes * slot);
    int k = 0;
    R* xk_d = buf + (size_t)chunk * k++;
    R* sj_d = buf + (size_t)chunk * k++;
    R* qg_d = buf + (size_t)chunk * k++;
    R* d_d = inverse ? buf + (size_t)chunk * k++ : nullptr;
    R* l_d = l_h ? buf + (size_t)chunk * k++ : nullptr;
    R* u_d = u_h ? buf + (size_t)chunk * k++ : nullptr;
    R* y_d = buf + (size_t)chunk * k++;
    SPX_CUDA(cudaMemcpyAsync(xk_d, xk_h + i0, bytes, cudaMemcpyHostToDevice, s));
    SPX_CUDA(cudaMemcpyAsync(sj_d, sj_h + i0, bytes, cudaMemcpyHostToDevice, s));
    SPX_CUDA(cudaMemcpyAsync(qg_d, qg_h + i0, bytes, cudaMemcpyHostToDevice, s));
    if (inverse) SPX_CUDA(cudaMemcpyAsync(d_d, d_h + i0, bytes, cudaMemcpyHostToDevice, s));
    if (l_h) SPX_CUDA(cudaMemcpyAsync(l_d, l_h + i0, bytes, cudaMemcpyHostToDevice, s));
    if (u_h) SPX_CUDA(cudaMemcpyAsync(u_d, u_h + i0, bytes, cudaMemcpyHostToDevice, s));
    Partial* part = ctx->d_partials + (size_t)slot * kMaxPartials;
    int nb = 0;
    st = launch_box<R>(ctx, s, op, inverse, m, y_d, xk_d, sj_d, qg_d, d_d, l_d, (R)l_val, u_d, (R)u_val, sel,
                       (R)lambda, (R)sigma, psi_out != nullptr, part, &nb, i0);
    if (st != SPX_OK) return st;
    if (psi_out) {
      st = enqueue_fold(ctx, s, part, nb, d_res + c);
      if (st != SPX_OK) return st;
    }
    SPX_CUDA(cudaMemcpyAsync(y_h + i0, y_d, bytes, cudaMemcpyDeviceToHost, s));
  }
  for (int i = 0; i < kSlots; ++i) SPX_CUDA(cudaStreamSynchronize(ctx->pipe_streams[i]));
  if (psi_out) {
    std::vector<Partial> res((size_t)nchunks);
    SPX_CUDA(cudaMemcpy(res.data(), d_res, res_bytes, cudaMemcpyDeviceToHost));
    double s = 0.0;
    bool bad = false;
    for (int64_t c = 0; c < nchunks; ++c) {
      s += res[(size_t)c].s;
      bad = bad || res[(size_t)c].bad > 0;
    }
    *psi_out = bad ? std::numeric_limits<double>::infinity() : (double)((R)lambda * (R)s);
  }
  return SPX_OK;
}

// Several Box operations at one shifted point, one pass over the host vectors.
template <class R, class Job>
static int32_t box_multi_host(spx_ctx* ctx, int32_t nops, const Job* jobs, int64_t n, const R* xk_h, const R* sj_h,
                              const R* l_h, double l_val, const R* u_h, double u_val, int64_t chunk,
                              double* psi_out) {
  SPX_REQUIRE(ctx != nullptr, "null context");
  SPX_REQUIRE(n >= 0, "n < 0");
  SPX_REQUIRE(nops >= 1 && nops <= 8 && jobs != nullptr, "nops must be 1..8");
  SPX_REQUIRE(n == 0 || (xk_h && sj_h), "null host vector");
  for (int j = 0; j < nops; ++j) {
    SPX_REQUIRE(jobs[j].op >= BOX_L1 && jobs[j].op <= BOX_LHALF, "unknown Box operator");
    SPX_REQUIRE(n == 0 || (jobs[j].y_host && jobs[j].q_or_g_host), "null host vector in job");
    SPX_REQUIRE(!(jobs[j].d_host != nullptr && jobs[j].op == BOX_LHALF), "RootNormLhalfBox has no iprox!");
  }
  DeviceGuard guard(ctx->device);
  if (n == 0) {
    if (psi_out) for (int j = 0; j < nops; ++j) psi_out[j] = 0.0;
    return SPX_OK;
  }
  if (chunk <= 0) chunk = (int64_t)1 << 22;
  chunk = (chunk + 3) & ~(int64_t)3;  // keeps every chunk pointer 16-byte aligned
  if (chunk > n) chunk = ((n + 3) & ~(int64_t)3);
  // distinct input vectors: xk, sj, [l], [u], then every q/g/d not seen before
  std::vector<const R*> uniq;
  auto slot_of = [&](const R* p) -> int {
    for (size_t k = 0; k < uniq.size(); ++k)
      if (uniq[k] == p) return (int)k;
    uniq.push_back(p);
    return (int)uniq.size() - 1;
  };
  const int i_xk = slot_of(xk_h), i_sj = slot_of(sj_h);
  const int i_l = l_h ? slot_of(l_h) : -1, i_u = u_h ? slot_of(u_h) : -1;
  int i_q[8], i_d[8];
  for (int j = 0; j < nops; ++j) {
    i_q[j] = slot_of(jobs[j].q_or_g_host);
    i_d[j] = jobs[j].d_host ? slot_of(jobs[j].d_host) : -1;
  }
  const int nin = (int)uniq.size();
  const size_t vec_bytes = (size_t)chunk * sizeof(R);
  const size_t slot_bytes = vec_bytes * (size_t)(nin + nops);
  const int64_t nchunks = (n + chunk - 1) / chunk;
  const size_t res_bytes = sizeof(Partial) * (size_t)nchunks * (size_t)nops;
  int32_t st = ensure_pipe(ctx, slot_bytes * kSlots + res_bytes + 256);
  if (st != SPX_OK) return st;
  char* base = (char*)ctx->pipe_buf;
  Partial* d_res = (Partial*)(base + slot_bytes * kSlots);
  DevSel sel;
  make_sel(nullptr, n, &sel);
  for (int64_t c = 0; c < nchunks; ++c) {
    const int slot = (int)(c % kSlots);
    cudaStream_t s = ctx->pipe_streams[slot];
    const int64_t i0 = c * chunk;
    const int64_t m = (n - i0) < chunk ? (n - i0) : chunk;
    const size_t bytes = (size_t)m * sizeof(R);
    R* buf = (R*)(base + slot_bytes * slot);
    auto dev = [&](int k) -> R* { return k < 0 ? nullptr : buf + (size_t)chunk * (size_t)k; };
    for (int k = 0; k < nin; ++k)
      SPX_CUDA(cudaMemcpyAsync(dev(k), uniq[(size_t)k] + i0, bytes, cudaMemcpyHostToDevice, s));
    for (int j = 0; j < nops; ++j) {
      R* y_d = dev(nin + j);
      // every op of a chunk runs on the same stream: the partial slots are reused in order
      Partial* part = ctx->d_partials + (size_t)slot * kMaxPartials;
      int nb = 0;
      st = launch_box<R>(ctx, s, jobs[j].op, jobs[j].d_host != nullptr, m, y_d, dev(i_xk), dev(i_sj), dev(i_q[j]),
                         dev(i_d[j]), dev(i_l), (R)l_val, dev(i_u), (R)u_val, sel, (R)jobs[j].lambda,
                         (R)jobs[j].sigma, psi_out != nullptr, part, &nb, i0);
      if (st != SPX_OK) return st;
      if (psi_out) {
        st = enqueue_fold(ctx, s, part, nb, d_res + c * nops + j);
        if (st != SPX_OK) return st;
      }
      SPX_CUDA(cudaMemcpyAsync(jobs[j].y_host + i0, y_d, bytes, cudaMemcpyDeviceToHost, s));
    }
  }
  for (int i = 0; i < kSlots; ++i) SPX_CUDA(cudaStreamSynchronize(ctx->pipe_streams[i]));
  if (psi_out) {
    std::vector<Partial> res((size_t)nchunks * (size_t)nops);
    SPX_CUDA(cudaMemcpy(res.data(), d_res, res_bytes, cudaMemcpyDeviceToHost));
    for (int j = 0; j < nops; ++j) {
      double sum = 0.0;
      bool bad = false;
      for (int64_t c = 0; c < nchunks; ++c) {
        sum += res[(size_t)(c * nops + j)].s;
        bad = bad || res[(size_t)(c * nops + j)].bad > 0;
      }
      psi_out[j] = bad ? std::numeric_limits<double>::infinity() : (double)((R)jobs[j].lambda * (R)sum);
    }
  }
  return SPX_OK;
}

}  // namespace spx

using namespace spx;

extern "C" int32_t spx_box_multi_host_f64(spx_ctx* ctx, int32_t nops, const spx_box_job_f64* jobs, int64_t n,
                                          const double* xk_host, const double* sj_host, const double* l_host,
                                          double l_val, const double* u_host, double u_val, int64_t chunk_elems,
                                          double* psi_out) {
  return box_multi_host<double>(ctx, nops, jobs, n, xk_host, sj_host, l_host, l_val, u_host, u_val, chunk_elems,
                                psi_out);
}
extern "C" int32_t spx_box_multi_host_f32(spx_ctx* ctx, int32_t nops, const spx_box_job_f32* jobs, int64_t n,
                                          const float* xk_host, const float* sj_host, const float* l_host,
                                          double l_val, const float* u_host, double u_val, int64_t chunk_elems,
                                          double* psi_out) {
  return box_multi_host<float>(ctx, nops, jobs, n, xk_host, sj_host, l_host, l_val, u_host, u_val, chunk_elems,
                               psi_out);
}

extern "C" int32_t spx_box_host_f64(spx_ctx* ctx, int32_t op, int64_t n, double* y_host, const double* xk_host,
                                    const double* sj_host, const double* q_or_g_host, const double* d_host,
                                    const double* l_host, double l_val, const double* u_host, double u_val,
                                    double lambda, double sigma, int64_t chunk_elems, double* psi_out) {
  return box_host<double>(ctx, op, n, y_host, xk_host, sj_host, q_or_g_host, d_host, l_host, l_val, u_host, u_val,
                          lambda, sigma, chunk_elems, psi_out);
}
extern "C" int32_t spx_box_host_f32(spx_ctx* ctx, int32_t op, int64_t n, float* y_host, const float* xk_host,
                                    const float* sj_host, const float* q_or_g_host, const float* d_host,
                                    const float* l_host, double l_val, const float* u_host, double u_val,
                                    double lambda, double sigma, int64_t chunk_elems, double* psi_out) {
  return box_host<float>(ctx, op, n, y_host, xk_host, sj_host, q_or_g_host, d_host, l_host, l_val, u_host, u_val,
                         lambda, sigma, chunk_elems, psi_out);
}
