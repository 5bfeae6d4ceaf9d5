// spx_comm.cu -- the collectives of the sharded path, inside the library and on the context's stream.
//
// One process per GPU; a vector (or a batch) is sharded contiguously over the ranks.  The only
// exchanges the path has are scalars: ψ(y) partial sums with their infeasibility flag
// (ShiftedProximalOperators.jl:51-54 and the Box / BInf overrides), the K partial sums of squares of every pass of
// the ℓ2 trust-region search (shiftedNormL1B2.jl:53-62) and the digit histograms of a single-vector top-r
// (shiftedIndBallL0.jl:66-70).  With a communicator attached to the context (spx_comm_init) and
// spx_comm_reduce_scalars(ctx, 1), every reduction the library folds on the device is all-reduced ON THE DEVICE
// (ncclAllReduce on ctx->stream, straight on the folded slots) before its one D2H copy: no host staging, no
// callback, no extra synchronisation.  Sums use ncclSum on Float64, flags ncclMax on Int64, so `Inf` survives
// as a flag (never Inf - Inf) and every rank receives the same bits.
//
// NCCL is bound at run time (dlopen of the libnccl.so.2 already mapped by the host framework, else the system
// one): the library itself has no link-time dependency and single-GPU users never load it.
#include <dlfcn.h>
#include <nccl.h>

#include <cstdlib>
#include <cstring>
#include <mutex>

#include "spx_common.cuh"

namespace spx {

struct NcclApi {
  void* handle = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*GroupStart)() = nullptr;
  ncclResult_t (*GroupEnd)() = nullptr;
  const char* (*GetErrorString)(ncclResult_t) = nullptr;
  ncclResult_t (*GetVersion)(int*) = nullptr;
};

static_assert((int)ncclInt64 == kNcclInt64 && (int)ncclUint64 == kNcclUint64 && (int)ncclFloat64 == kNcclFloat64 &&
                  (int)ncclSum == kNcclSum && (int)ncclMax == kNcclMax,
              "spx_common.cuh NCCL constants");
static NcclApi g_nccl;
static std::once_flag g_nccl_once;

static void load_nccl() {
  const char* names[4] = {std::getenv("SPX_NCCL_LIB"), "libnccl.so.2", "libnccl.so", nullptr};
  void* h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD);  // the copy the host framework already mapped
  for (int i = 0; h == nullptr && i < 3; ++i)
    if (names[i] != nullptr && names[i][0] != 0) h = dlopen(names[i], RTLD_NOW | RTLD_GLOBAL);
  if (h == nullptr) return;
  NcclApi a;
  a.handle = h;
  a.GetUniqueId = (decltype(a.GetUniqueId))dlsym(h, "ncclGetUniqueId");
  a.CommInitRank = (decltype(a.CommInitRank))dlsym(h, "ncclCommInitRank");
  a.CommDestroy = (decltype(a.CommDestroy))dlsym(h, "ncclCommDestroy");
  a.AllReduce = (decltype(a.AllReduce))dlsym(h, "ncclAllReduce");
  a.GroupStart = (decltype(a.GroupStart))dlsym(h, "ncclGroupStart");
  a.GroupEnd = (decltype(a.GroupEnd))dlsym(h, "ncclGroupEnd");
  a.GetErrorString = (decltype(a.GetErrorString))dlsym(h, "ncclGetErrorString");
  a.GetVersion = (decltype(a.GetVersion))dlsym(h, "ncclGetVersion");
  if (a.GetUniqueId && a.CommInitRank && a.CommDestroy && a.AllReduce && a.GroupStart && a.GroupEnd) g_nccl = a;
}

static const NcclApi* nccl() {
  std::call_once(g_nccl_once, load_nccl);
  return g_nccl.handle ? &g_nccl : nullptr;
}

static int32_t nccl_fail(ncclResult_t r, const char* what) {
  const NcclApi* a = nccl();
  set_error("NCCL error %d (%s) in %s", (int)r, (a && a->GetErrorString) ? a->GetErrorString(r) : "?", what);
  return 1000 + (int32_t)r;
}
#define SPX_NCCL(call)                                        \
  do {                                                        \
    ncclResult_t r__ = (call);                                \
    if (r__ != ncclSuccess) return spx::nccl_fail(r__, #call); \
  } while (0)

// slots of Partial {s, s2, bad} -> dbuf[0..nslot) = s, dbuf[nslot..2 nslot) = s2, ibuf[0..nslot) = bad, and back
__global__ void pack_partials_kernel(const Partial* __restrict__ r, int nslot, double* __restrict__ dbuf,
                                     long long* __restrict__ ibuf) {
  const int k = threadIdx.x;
  if (k < nslot) {
    dbuf[k] = r[k].s;
    dbuf[nslot + k] = r[k].s2;
    ibuf[k] = r[k].bad;
  }
}
__global__ void unpack_partials_kernel(Partial* __restrict__ r, int nslot, const double* __restrict__ dbuf,
                                       const long long* __restrict__ ibuf) {
  const int k = threadIdx.x;
  if (k < nslot) {
    r[k].s = dbuf[k];
    r[k].s2 = dbuf[nslot + k];
    r[k].bad = ibuf[k];
  }
}
__global__ void neutral_partials_kernel(Partial* __restrict__ r, int nslot) {
  const int k = threadIdx.x;
  if (k < nslot) {
    r[k].s = 0.0;
    r[k].s2 = 0.0;
    r[k].bad = -1;
  }
}

bool comm_active(const spx_ctx* ctx) {
  return (ctx->comm != nullptr || ctx->d_peer_ptrs != nullptr) && ctx->reduce_scalars && ctx->comm_nranks > 1;
}

int32_t comm_neutral_result(spx_ctx* ctx, int nslot) {
  neutral_partials_kernel<<<1, kMaxScale, 0, ctx->stream>>>(ctx->d_result, nslot);
  ctx->launches++;
  SPX_CUDA(cudaGetLastError());
  return SPX_OK;
}

// All-reduce of ctx->d_result[0..nslot) over the ranks, enqueued on ctx->stream: Σ (s, s2) and max (bad).
int32_t comm_allreduce_result(spx_ctx* ctx, int nslot) {
  const NcclApi* a = nccl();
  if (a == nullptr || ctx->comm == nullptr) {
    set_error("no communicator on this context (spx_comm_init)");
    return SPX_E_INVALID;
  }
  ncclComm_t comm = (ncclComm_t)ctx->comm;
  if (nslot == 1) {  // {s, s2} are two adjacent doubles, bad one int64: in place, one grouped launch
    SPX_NCCL(a->GroupStart());
    SPX_NCCL(a->AllReduce(&ctx->d_result[0].s, &ctx->d_result[0].s, 2, ncclDouble, ncclSum, comm, ctx->stream));
    SPX_NCCL(a->AllReduce(&ctx->d_result[0].bad, &ctx->d_result[0].bad, 1, ncclInt64, ncclMax, comm, ctx->stream));
    SPX_NCCL(a->GroupEnd());
    ctx->collectives += 1;
    return SPX_OK;
  }
  double* dbuf = ctx->d_comm;
  long long* ibuf = (long long*)(ctx->d_comm + 2 * kMaxScale);
  pack_partials_kernel<<<1, kMaxScale, 0, ctx->stream>>>(ctx->d_result, nslot, dbuf, ibuf);
  SPX_NCCL(a->GroupStart());
  SPX_NCCL(a->AllReduce(dbuf, dbuf, 2 * (size_t)nslot, ncclDouble, ncclSum, comm, ctx->stream));
  SPX_NCCL(a->AllReduce(ibuf, ibuf, (size_t)nslot, ncclInt64, ncclMax, comm, ctx->stream));
  SPX_NCCL(a->GroupEnd());
  unpack_partials_kernel<<<1, kMaxScale, 0, ctx->stream>>>(ctx->d_result, nslot, dbuf, ibuf);
  ctx->launches += 2;
  ctx->collectives += 1;
  SPX_CUDA(cudaGetLastError());
  return SPX_OK;
}

int32_t comm_allreduce_raw(spx_ctx* ctx, void* buf, size_t count, int dtype, int op) {
  const NcclApi* a = nccl();
  if (a == nullptr || ctx->comm == nullptr) {
    set_error("no communicator on this context (spx_comm_init)");
    return SPX_E_INVALID;
  }
  SPX_NCCL(a->AllReduce(buf, buf, count, (ncclDataType_t)dtype, (ncclRedOp_t)op, (ncclComm_t)ctx->comm, ctx->stream));
  ctx->collectives += 1;
  return SPX_OK;
}

// ------------------------------------------------------------------ peer-memory all-reduce ---
// The reductions of the path are 1..16 slots of {Σ, Σ₂, flag}: latency, not bandwidth.  With every rank's exchange
// buffer mapped into every process (CUDA IPC over NVLink / NVSwitch) the fold kernel finishes the all-reduce itself:
//   block k folds slot k of this rank's partials; threads 0..nranks-1 store the folded slot -- data, a system-scope
//   fence, then the sequence number -- into bank (seq & 1), row `rank`, of EVERY rank's buffer; the same threads then
//   spin (acquire loads) until row w of the OWN buffer carries this sequence number; thread 0 sums the rows in rank
//   order, so every rank computes the same bits; the flag is max-reduced.
// Two banks suffice: a rank can only be one reduction ahead of a peer (it needs the peer's row to finish its own).
// A rank that never arrives would hang the others exactly like an NCCL collective; the spin is bounded by the
// global timer (SPX_PEER_TIMEOUT_S, default 120 s) and poisons the flag instead, which the host turns into an error.
struct PeerSlot {
  double s, s2;
  long long bad;
  unsigned long long seq;
};
constexpr int kPeerMaxRanks = 16;
constexpr size_t kPeerBytes = sizeof(PeerSlot) * 2 * kPeerMaxRanks * kMaxScale;

__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ unsigned long long globaltimer_ns() {
  unsigned long long v;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(v));
  return v;
}
__device__ __forceinline__ void st_release_sys(unsigned long long* p, unsigned long long v) {
  asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}

__global__ void __launch_bounds__(256) fold_allreduce_peer_kernel(const Partial* __restrict__ partials, int nblocks,
                                                                  Partial* __restrict__ result, void* const* peers,
                                                                  int nranks, int rank, unsigned long long seq,
                                                                  unsigned long long timeout_ns) {
  const int k = blockIdx.x;  // slot
  Partial acc;
  acc.s = 0.0;
  acc.s2 = 0.0;
  acc.bad = -1;
  const Partial* p = partials + (size_t)k * nblocks;
  for (int i = threadIdx.x; i < nblocks; i += 256) {
    Partial t = p[i];
    acc.s += t.s;
    acc.s2 += t.s2;
    acc.bad = t.bad > acc.bad ? t.bad : acc.bad;
  }
  acc = block_fold<256>(acc);
  __shared__ Partial mine;
  __shared__ int timed_out;
  if (threadIdx.x == 0) {
    mine = acc;
    timed_out = 0;
  }
  __syncthreads();
  const int bank = (int)(seq & 1ull);
  const int t = threadIdx.x;
  if (t < nranks) {  // push: my folded slot into row `rank` of peer t's buffer
    PeerSlot* dst = (PeerSlot*)peers[t] + ((size_t)bank * kPeerMaxRanks + rank) * kMaxScale + k;
    dst->s = mine.s;
    dst->s2 = mine.s2;
    dst->bad = mine.bad;
    __threadfence_system();
    st_release_sys(&dst->seq, seq);
    // pull: wait for row t of my own buffer
    const PeerSlot* src = (const PeerSlot*)peers[rank] + ((size_t)bank * kPeerMaxRanks + t) * kMaxScale + k;
    unsigned long long t0 = 0;
    unsigned spins = 0;
    while (ld_acquire_sys(&src->seq) != seq) {
      if ((++spins & 1023u) == 0) {  // look at the clock every ~1000 polls
        const unsigned long long now = globaltimer_ns();
        if (t0 == 0) t0 = now;
        if (now - t0 > timeout_ns) {
          timed_out = 1;
          break;
        }
      }
      __nanosleep(32);
    }
  }
  __syncthreads();
  if (t == 0) {
    Partial out;
    out.s = 0.0;
    out.s2 = 0.0;
    out.bad = -1;
    const PeerSlot* base = (const PeerSlot*)peers[rank] + (size_t)bank * kPeerMaxRanks * kMaxScale + k;
    for (int w = 0; w < nranks; ++w) {  // rank order: the same sum on every rank
      const volatile PeerSlot* v = base + (size_t)w * kMaxScale;  // written by a peer during this kernel: no L1
      const double vs = v->s, vs2 = v->s2;
      const long long vb = v->bad;
      out.s += vs;
      out.s2 += vs2;
      out.bad = vb > out.bad ? vb : out.bad;
    }
    if (timed_out) out.bad = 1ll << 61;
    result[k] = out;
  }
}

static unsigned long long peer_timeout_ns() {  // SPX_PEER_TIMEOUT_S: how long a rank waits for its peers (default 120 s)
  static const unsigned long long v = [] {
    const char* e = getenv("SPX_PEER_TIMEOUT_S");
    double sec = e ? atof(e) : 120.0;
    if (!(sec > 0.0)) sec = 120.0;
    return (unsigned long long)(sec * 1e9);
  }();
  return v;
}

bool comm_peer_ready(const spx_ctx* ctx) { return ctx->d_peer_ptrs != nullptr && ctx->peer_nranks == ctx->comm_nranks; }

int32_t comm_fold_allreduce_peer(spx_ctx* ctx, int nblocks, int nslot) {
  ctx->peer_seq += 1;
  fold_allreduce_peer_kernel<<<nslot, 256, 0, ctx->stream>>>(ctx->d_partials, nblocks > 0 ? nblocks : 0, ctx->d_result,
                                                             ctx->d_peer_ptrs, ctx->peer_nranks, ctx->comm_rank,
                                                             ctx->peer_seq, peer_timeout_ns());
  ctx->launches++;
  ctx->collectives += 1;
  SPX_CUDA(cudaGetLastError());
  return SPX_OK;
}

}  // namespace spx

using namespace spx;

extern "C" {

int32_t spx_comm_available(void) { return nccl() != nullptr ? 1 : 0; }

int32_t spx_comm_unique_id(void* id_out) {
  SPX_REQUIRE(id_out != nullptr, "null id");
  const NcclApi* a = nccl();
  if (a == nullptr) {
    set_error("libnccl.so.2 not found (set SPX_NCCL_LIB)");
    return SPX_E_UNSUPPORTED;
  }
  static_assert(sizeof(ncclUniqueId) == SPX_COMM_ID_BYTES, "unique id size");
  SPX_NCCL(a->GetUniqueId((ncclUniqueId*)id_out));
  return SPX_OK;
}

int32_t spx_comm_init(spx_ctx* ctx, int32_t nranks, int32_t rank, const void* id) {
  SPX_REQUIRE(ctx != nullptr && id != nullptr, "null argument");
  SPX_REQUIRE(nranks >= 1 && rank >= 0 && rank < nranks, "bad rank / nranks");
  SPX_REQUIRE(ctx->comm == nullptr, "the context already has a communicator");
  const NcclApi* a = nccl();
  if (a == nullptr) {
    set_error("libnccl.so.2 not found (set SPX_NCCL_LIB)");
    return SPX_E_UNSUPPORTED;
  }
  DeviceGuard g(ctx->device);
  ncclUniqueId uid;
  memcpy(&uid, id, sizeof(uid));
  ncclComm_t comm = nullptr;
  SPX_NCCL(a->CommInitRank(&comm, nranks, uid, rank));
  if (ctx->d_comm == nullptr) SPX_CUDA(cudaMalloc(&ctx->d_comm, sizeof(double) * 3 * kMaxScale + 64));
  ctx->comm = comm;
  ctx->comm_nranks = nranks;
  ctx->comm_rank = rank;
  return SPX_OK;
}

int32_t spx_comm_destroy(spx_ctx* ctx) {
  SPX_REQUIRE(ctx != nullptr, "null context");
  if (ctx->comm != nullptr) {
    const NcclApi* a = nccl();
    DeviceGuard g(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    if (a != nullptr) a->CommDestroy((ncclComm_t)ctx->comm);
    ctx->comm = nullptr;
  }
  ctx->comm_nranks = 1;
  ctx->comm_rank = 0;
  ctx->reduce_scalars = false;
  return SPX_OK;
}

int32_t spx_comm_reduce_scalars(spx_ctx* ctx, int32_t on) {
  SPX_REQUIRE(ctx != nullptr, "null context");
  SPX_REQUIRE(on == 0 || ctx->comm != nullptr || ctx->d_peer_ptrs != nullptr,
              "no communicator on this context (spx_comm_init / spx_comm_peer_attach)");
  ctx->reduce_scalars = on != 0;
  return SPX_OK;
}

int32_t spx_comm_info(spx_ctx* ctx, int32_t* nranks, int32_t* rank, int64_t* collectives) {
  SPX_REQUIRE(ctx != nullptr, "null context");
  if (nranks) *nranks = ctx->comm ? ctx->comm_nranks : 1;
  if (rank) *rank = ctx->comm ? ctx->comm_rank : 0;
  if (collectives) *collectives = ctx->collectives;
  return SPX_OK;
}

int32_t spx_comm_peer_export(spx_ctx* ctx, void* handle_out) {
  SPX_REQUIRE(ctx != nullptr && handle_out != nullptr, "null argument");
  static_assert(sizeof(cudaIpcMemHandle_t) == SPX_PEER_HANDLE_BYTES, "IPC handle size");
  DeviceGuard g(ctx->device);
  if (ctx->peer_own == nullptr) {
    SPX_CUDA(cudaMalloc(&ctx->peer_own, kPeerBytes));
    SPX_CUDA(cudaMemset(ctx->peer_own, 0, kPeerBytes));
  }
  cudaIpcMemHandle_t h;
  SPX_CUDA(cudaIpcGetMemHandle(&h, ctx->peer_own));
  memcpy(handle_out, &h, sizeof(h));
  return SPX_OK;
}

int32_t spx_comm_peer_attach(spx_ctx* ctx, int32_t nranks, int32_t rank, const void* handles) {
  SPX_REQUIRE(ctx != nullptr && handles != nullptr, "null argument");
  SPX_REQUIRE(nranks >= 1 && nranks <= kPeerMaxRanks && rank >= 0 && rank < nranks, "bad rank / nranks");
  SPX_REQUIRE(ctx->peer_own != nullptr, "call spx_comm_peer_export first");
  SPX_REQUIRE(ctx->d_peer_ptrs == nullptr, "exchange buffers already attached");
  DeviceGuard g(ctx->device);
  void* ptrs[kPeerMaxRanks] = {};
  for (int r = 0; r < nranks; ++r) {
    if (r == rank) {
      ptrs[r] = ctx->peer_own;
      continue;
    }
    cudaIpcMemHandle_t h;
    memcpy(&h, (const char*)handles + (size_t)r * sizeof(h), sizeof(h));
    cudaError_t e = cudaIpcOpenMemHandle(&ptrs[r], h, cudaIpcMemLazyEnablePeerAccess);
    if (e != cudaSuccess) {
      for (int q = 0; q < r; ++q)
        if (q != rank && ptrs[q]) cudaIpcCloseMemHandle(ptrs[q]);
      return cuda_fail(e, "cudaIpcOpenMemHandle (peer exchange buffer)");
    }
    ctx->peer_mapped[r] = ptrs[r];
  }
  SPX_CUDA(cudaMalloc((void**)&ctx->d_peer_ptrs, sizeof(void*) * kPeerMaxRanks));
  SPX_CUDA(cudaMemcpy(ctx->d_peer_ptrs, ptrs, sizeof(void*) * kPeerMaxRanks, cudaMemcpyHostToDevice));
  ctx->peer_nranks = nranks;
  ctx->comm_rank = rank;
  if (ctx->comm == nullptr) ctx->comm_nranks = nranks;  // the peer exchange alone carries the scalar reductions
  ctx->peer_seq = 0;
  return SPX_OK;
}

int32_t spx_comm_peer_active(spx_ctx* ctx) { return ctx != nullptr && comm_peer_ready(ctx) ? 1 : 0; }

int32_t spx_comm_peer_detach(spx_ctx* ctx) {
  SPX_REQUIRE(ctx != nullptr, "null context");
  DeviceGuard g(ctx->device);
  cudaStreamSynchronize(ctx->stream);
  for (int r = 0; r < kPeerMaxRanks; ++r)
    if (ctx->peer_mapped[r]) {
      cudaIpcCloseMemHandle(ctx->peer_mapped[r]);
      ctx->peer_mapped[r] = nullptr;
    }
  if (ctx->d_peer_ptrs) cudaFree(ctx->d_peer_ptrs);
  ctx->d_peer_ptrs = nullptr;
  ctx->peer_nranks = 0;
  return SPX_OK;
}

int32_t spx_comm_allreduce_f64(spx_ctx* ctx, double* dev_buf, int64_t count, int32_t op) {
  SPX_REQUIRE(ctx != nullptr && dev_buf != nullptr && count >= 0, "bad argument");
  SPX_REQUIRE(op == 0 || op == 1, "op: 0 sum, 1 max");
  if (ctx->comm == nullptr || ctx->comm_nranks == 1 || count == 0) return SPX_OK;
  DeviceGuard g(ctx->device);
  return comm_allreduce_raw(ctx, dev_buf, (size_t)count, (int)ncclDouble, op == 0 ? (int)ncclSum : (int)ncclMax);
}

}  // extern "C"
