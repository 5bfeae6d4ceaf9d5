// spx_topr.cu -- ShiftedIndBallL0 / ShiftedIndBallL0BInf prox!: on-device radix-select.
//
// shiftedIndBallL0.jl:54-72, shiftedIndBallL0BInf.jl:73-95.  The reference
// sorts a permutation (`sortperm!(p, y, rev=true, by=abs)`) and zeroes all but
// the first r entries; equal |z| are ordered by ascending index (Base.Perm
// ordering), NaN sorts largest.  Here no permutation exists:
//
//   one thread-block CLUSTER owns one problem.  Each CTA streams its slice of
//   xk, sj, q from HBM exactly once (128-bit loads), keeps z = (xk+sj)+q in
//   registers (16 per thread) and xk+sj in shared memory, and the cluster runs
//   an MSD radix-select (11-bit digits) on the order-preserving integer image
//   of |z|: per-CTA histograms in shared memory (warp-aggregated atomics),
//   summed across the cluster through distributed shared memory.  Ties at the
//   threshold are resolved by an ordered prefix count (lowest index first).
//   y is written once.  HBM traffic = the algorithmic 4R per element.
//
// Problems longer than 8 x 16384 elements take the multi-pass global path at
// the end of this file (z stashed in y, one histogram pass per digit).
#include <cooperative_groups.h>

#include <cstring>
#include <vector>

#include "spx_common.cuh"

namespace cg = cooperative_groups;

namespace spx {

#ifndef SPX_TR_THREADS
#define SPX_TR_THREADS 1024
#endif
constexpr int kTrThreads = SPX_TR_THREADS;     // 1024 x 64 registers: one CTA per SM (512 x 2 measured slower: 8-CTA clusters)
constexpr int kTrE = 16;                       // elements per thread (registers)
constexpr int kTrChunk = kTrThreads * kTrE;    // 16384 elements per CTA
constexpr int kTrBins = 2048;                  // 11-bit digits
constexpr int kTrBpt = kTrBins / kTrThreads;   // bins per thread when a digit is picked
constexpr int kTrMaxCluster = 8;
constexpr int kTrCandMax = 1024;               // candidates of the threshold bin resolved by direct ranking
constexpr int kPickThreads = 1024;             // single-block digit pick of the global path

#ifdef SPX_TR_TIMING
__device__ unsigned long long g_tr_t[16];
#define TR_T(i)                                          \
  do {                                                   \
    if (threadIdx.x == 0 && blockIdx.x == 0) {           \
      const long long now__ = clock64();                 \
      atomicAdd(&g_tr_t[i], (unsigned long long)(now__ - tr_last)); \
      tr_last = now__;                                   \
    }                                                    \
  } while (0)
#else
#define TR_T(i)
#endif

template <class R> struct KeyTraits;
template <> struct KeyTraits<double> {
  using K = unsigned long long;
  static constexpr int BITS = 63;
  static constexpr int NPASS = 6;
  __device__ static __forceinline__ K key(double z) {
    K k = (K)__double_as_longlong(z) & 0x7fffffffffffffffull;
    return k > 0x7ff0000000000000ull ? 0x7fffffffffffffffull : k;  // every NaN is "largest"
  }
};
template <> struct KeyTraits<float> {
  using K = unsigned int;
  static constexpr int BITS = 31;
  static constexpr int NPASS = 3;
  __device__ static __forceinline__ K key(float z) {
    K k = (K)__float_as_int(z) & 0x7fffffffu;
    return k > 0x7f800000u ? 0x7fffffffu : k;
  }
};
// digit widths, most significant first
__host__ __device__ constexpr int digit_bits(int BITS, int pass) {
  return (BITS - 11 * pass) >= 11 ? 11 : (BITS - 11 * pass);
}

// exclusive prefix sum of one int per thread over the block (ascending thread id);
// `ws` holds 33 ints.  Returns the exclusive prefix; *total = block sum.
template <int THREADS> __device__ __forceinline__ int block_excl_scan(int v, int* ws, int* total) {
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  int inc = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    int t = __shfl_up_sync(0xffffffffu, inc, o);
    if (lane >= o) inc += t;
  }
  __syncthreads();  // ws reuse
  if (lane == 31) ws[w] = inc;
  __syncthreads();
  if (w == 0) {
    int s = lane < (THREADS / 32) ? ws[lane] : 0;
    int sinc = s;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      int t = __shfl_up_sync(0xffffffffu, sinc, o);
      if (lane >= o) sinc += t;
    }
    ws[lane] = sinc - s;  // exclusive warp offsets
    if (lane == 31) ws[32] = sinc;
  }
  __syncthreads();
  *total = ws[32];
  return ws[w] + inc - v;
}

struct TrShared {
  unsigned hist[kTrBins];
  unsigned tot[kTrBins];
  int ws[40];
  int sel_bin;
  long long sel_above;
  long long sel_count;
  long long eq_count;  // equal-to-threshold elements in this CTA
  // linear-bin form
  unsigned long long kmax_slot[kTrMaxCluster];  // every CTA's largest key, replicated in every CTA
  unsigned long long cand_key[kTrCandMax];      // CTA 0: keys of the threshold bin
  int cand_idx[kTrCandMax];                     //        and their positions in the problem
  int cand_count;                               // CTA 0
  unsigned long long thr_key;                   // replicated: the r-th largest key ...
  int thr_idx;                                  // ... and the last position kept among its equals
};

template <class R, bool BINF, bool VECLD>
__global__ void __launch_bounds__(kTrThreads, 65536 / (64 * kTrThreads))
    topr_cluster_kernel(R* y, const R* xk, const R* sj, const R* q, long long n, long long r, R delta,
                        long long nprob, const unsigned char* __restrict__ only_flagged) {
  using KT = KeyTraits<R>;
  using K = typename KT::K;
  constexpr int VEC = VECLD ? 16 / (int)sizeof(R) : 1;
  constexpr int ROUNDS = kTrE / VEC;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  TrShared* sh = reinterpret_cast<TrShared*>(smem_raw);
  R* xs_sm = reinterpret_cast<R*>(smem_raw + ((sizeof(TrShared) + 15) / 16) * 16);

  cg::cluster_group cluster = cg::this_cluster();
  const unsigned csize = cluster.num_blocks();
  const unsigned crank = cluster.block_rank();
  const long long ncluster = gridDim.x / csize;
  const long long cid = blockIdx.x / csize;
  const int t = threadIdx.x;
  const int lane = t & 31;

  for (long long prob = cid; prob < nprob; prob += ncluster) {
    if (only_flagged != nullptr && only_flagged[prob] == 0) continue;  // cluster-uniform
    const long long pbase = prob * n;
    const long long cbase = (long long)crank * kTrChunk;  // this CTA's slice of the problem
    long long cnt = n - cbase;
    cnt = cnt < 0 ? 0 : (cnt > kTrChunk ? kTrChunk : cnt);

#ifdef SPX_TR_TIMING
    long long tr_last = clock64();
#endif
    // ---- load: z in registers, xs in shared memory --------------------------
    R z[kTrE];
    unsigned valid = 0;
#pragma unroll
    for (int k = 0; k < ROUNDS; ++k) {
      const long long li = ((long long)k * kTrThreads + t) * VEC;
      if (li < cnt) {  // VECLD => cnt is a multiple of VEC
        Pack<R, VEC> a, b, c;
        ld_stream(xk + pbase + cbase + li, a);
        ld_stream(sj + pbase + cbase + li, b);
        ld_stream(q + pbase + cbase + li, c);
#pragma unroll
        for (int e = 0; e < VEC; ++e) {
          const R xs = a.v[e] + b.v[e];
          z[k * VEC + e] = xs + c.v[e];  // (xk + sj) + q   shiftedIndBallL0.jl:66
          xs_sm[li + e] = xs;
          valid |= 1u << (k * VEC + e);
        }
      } else {
#pragma unroll
        for (int e = 0; e < VEC; ++e) z[k * VEC + e] = R(0);
      }
    }

    __syncthreads();
    TR_T(0);  // load
    // ---- radix select of the r-th largest key ------------------------------
    bool keep_all_bin = true;   // every key whose prefix equals `prefix` is kept
    K prefix = 0;               // selected digits so far
    int shift = KT::BITS;       // bits below the known prefix
    long long need = r;         // rank still to resolve inside the prefix bin
    const bool select = (r > 0) && (r < n);
    if (select) {
      keep_all_bin = false;
      unsigned cand = valid;  // slots whose key still carries the selected prefix
#pragma unroll 1
      for (int pass = 0; pass < KT::NPASS; ++pass) {
        const int width = digit_bits(KT::BITS, pass);
        shift -= width;
        for (int b = t; b < kTrBins; b += kTrThreads) sh->hist[b] = 0;
        __syncthreads();
        const K dmask = (K)((1u << width) - 1u);
        if (__any_sync(0xffffffffu, cand != 0u)) {  // after the first digits most warps hold no candidate
#pragma unroll
          for (int s = 0; s < kTrE; ++s) {
            const bool in = (cand >> s) & 1u;
            if (pass > 0 && !__any_sync(0xffffffffu, in)) continue;
            const unsigned d = in ? (unsigned)((KT::key(z[s]) >> shift) & dmask) : 0xffffffffu;
            const unsigned peers = __match_any_sync(0xffffffffu, d);
            if (in && lane == (__ffs(peers) - 1)) atomicAdd(&sh->hist[d], (unsigned)__popc(peers));
          }
        }
        __syncthreads();
        TR_T(1);  // local histogram
        cluster.sync();  // every CTA's histogram is complete
        TR_T(2);  // cluster sync
        for (int b = t; b < kTrBins; b += kTrThreads) {
          unsigned s = 0;
          for (unsigned rr = 0; rr < csize; ++rr) s += cluster.map_shared_rank(sh->hist, rr)[b];
          sh->tot[b] = s;
        }
        cluster.sync();  // remote reads done before anyone clears its histogram again
        TR_T(3);  // remote sum + sync
        // bins in descending order: thread t owns bins 2047 - kTrBpt t ... 2048 - kTrBpt (t + 1)
        int c[kTrBpt], mine = 0;
#pragma unroll
        for (int j = 0; j < kTrBpt; ++j) {
          c[j] = (int)sh->tot[kTrBins - 1 - kTrBpt * t - j];
          mine += c[j];
        }
        int total;
        long long above = block_excl_scan<kTrThreads>(mine, sh->ws, &total);
#pragma unroll
        for (int j = 0; j < kTrBpt; ++j) {
          if (above < need && need <= above + c[j]) {
            sh->sel_bin = kTrBins - 1 - kTrBpt * t - j;
            sh->sel_above = above;
            sh->sel_count = c[j];
          }
          above += c[j];
        }
        __syncthreads();
        const unsigned sel = (unsigned)sh->sel_bin;
        prefix = (prefix << width) | (K)sel;
        need -= sh->sel_above;
        const long long bin_count = sh->sel_count;
        __syncthreads();
        unsigned keepc = 0;
#pragma unroll
        for (int s = 0; s < kTrE; ++s)
          if (((cand >> s) & 1u) && (unsigned)((KT::key(z[s]) >> shift) & dmask) == sel) keepc |= 1u << s;
        cand = keepc;
        TR_T(4);  // scan + pick
        if (need == bin_count) {  // the whole bin is kept: no finer digit needed
          keep_all_bin = true;
          break;
        }
      }
    }
    // after the loop: element kept iff (key >> shift) > prefix, or == prefix and
    // (keep_all_bin or its rank among the equal elements, in index order, < need)
    const bool keep_everything = !select && r >= n;
    const bool keep_nothing = !select && r <= 0;

    unsigned keepmask = 0;
    if (keep_everything) {
      keepmask = valid;
    } else if (!keep_nothing) {
      unsigned eqmask = 0;
#pragma unroll
      for (int s = 0; s < kTrE; ++s) {
        if ((valid >> s) & 1u) {
          const K kp = KT::key(z[s]) >> shift;
          if (kp > prefix) keepmask |= 1u << s;
          else if (kp == prefix) eqmask |= 1u << s;
        }
      }
      if (keep_all_bin) {
        keepmask |= eqmask;
      } else {
        // ordered tie resolution: equal elements ranked by global index
        int total;
        int mine = __popc(eqmask);
        (void)block_excl_scan<kTrThreads>(mine, sh->ws, &total);
        if (t == 0) sh->eq_count = total;
        cluster.sync();
        long long offset = 0;
        for (unsigned rr = 0; rr < crank; ++rr) offset += *cluster.map_shared_rank(&sh->eq_count, rr);
        cluster.sync();
#pragma unroll 1
        for (int k = 0; k < ROUNDS; ++k) {
          const unsigned m = (eqmask >> (k * VEC)) & ((1u << VEC) - 1u);
          int rtotal;
          long long rank = offset + block_excl_scan<kTrThreads>(__popc(m), sh->ws, &rtotal);
#pragma unroll
          for (int e = 0; e < VEC; ++e) {
            if ((m >> e) & 1u) {
              if (rank < need) keepmask |= 1u << (k * VEC + e);
              ++rank;
            }
          }
          offset += rtotal;
        }
      }
    }

    __syncthreads();
    TR_T(5);  // keep masks / ties
    // ---- write y once -------------------------------------------------------
#pragma unroll
    for (int k = 0; k < ROUNDS; ++k) {
      const long long li = ((long long)k * kTrThreads + t) * VEC;
      if (li < cnt) {
        Pack<R, VEC> o;
#pragma unroll
        for (int e = 0; e < VEC; ++e) {
          const int s = k * VEC + e;
          const R zz = ((keepmask >> s) & 1u) ? z[s] : R(0);
          R v = zz - xs_sm[li + e];  // y .-= xk .+ sj   (:70)
          if (BINF) v = jl_min(jl_max(v, -delta), delta);  // shiftedIndBallL0BInf.jl:91
          o.v[e] = v;
        }
        st_stream(y + pbase + cbase + li, o);
      }
    }
    cluster.sync();  // shared state is reused by the next problem
    TR_T(6);  // write + final sync
  }
}

// ------------------------------------------------------ linear-bin form ---
// The radix digits of an IEEE key are a poor first cut: the top 11 bits are the exponent, which a whole
// problem shares up to a handful of values, so the first pass separates nothing and the select needs ~3
// passes of warp-matched histogram atomics and ~8 cluster barriers.  This form bins |z| LINEARLY between 0
// and the problem's largest magnitude (2048 bins, monotone in |z|, spread over the whole range, so plain
// shared-memory atomics rarely collide): one histogram pass locates the bin holding the r-th largest
// element; that bin holds ~n/2048 elements, which are gathered into CTA 0 and ranked directly (key
// descending, position ascending -- the reference's tie order), giving the exact threshold (key, position).
// Problems it cannot decide cheaply -- a NaN or Inf, all zeros, more than 1024 candidates in the threshold
// bin -- are flagged and left to the radix kernel above, which runs afterwards on the flagged problems only.
template <class R, bool BINF, bool VECLD>
__global__ void __launch_bounds__(kTrThreads, 1)
    topr_cluster_lin_kernel(R* y, const R* xk, const R* sj, const R* q, long long n, long long r, R delta,
                            long long nprob, unsigned char* __restrict__ fallback) {
  using KT = KeyTraits<R>;
  using K = typename KT::K;
  constexpr int VEC = VECLD ? 16 / (int)sizeof(R) : 1;
  constexpr int ROUNDS = kTrE / VEC;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  TrShared* sh = reinterpret_cast<TrShared*>(smem_raw);
  R* xs_sm = reinterpret_cast<R*>(smem_raw + ((sizeof(TrShared) + 15) / 16) * 16);

  cg::cluster_group cluster = cg::this_cluster();
  const unsigned csize = cluster.num_blocks();
  const unsigned crank = cluster.block_rank();
  const long long ncluster = gridDim.x / csize;
  const long long cid = blockIdx.x / csize;
  const int t = threadIdx.x;
  const int lane = t & 31;
  const bool select = (r > 0) && (r < n);
  TrShared* sh0 = cluster.map_shared_rank(sh, 0);

  for (long long prob = cid; prob < nprob; prob += ncluster) {
    const long long pbase = prob * n;
    const long long cbase = (long long)crank * kTrChunk;  // this CTA's slice of the problem
    long long cnt = n - cbase;
    cnt = cnt < 0 ? 0 : (cnt > kTrChunk ? kTrChunk : cnt);

#ifdef SPX_TR_TIMING
    long long tr_last = clock64();
#endif
    // ---- load: z in registers, xs in shared memory --------------------------
    R z[kTrE];
    unsigned valid = 0;
    K kmax = 0;
#pragma unroll
    for (int k = 0; k < ROUNDS; ++k) {
      const long long li = ((long long)k * kTrThreads + t) * VEC;
      if (li < cnt) {  // VECLD => cnt is a multiple of VEC
        Pack<R, VEC> a, b, c;
        ld_stream(xk + pbase + cbase + li, a);
        ld_stream(sj + pbase + cbase + li, b);
        ld_stream(q + pbase + cbase + li, c);
#pragma unroll
        for (int e = 0; e < VEC; ++e) {
          const R xs = a.v[e] + b.v[e];
          z[k * VEC + e] = xs + c.v[e];  // (xk + sj) + q   shiftedIndBallL0.jl:66
          xs_sm[li + e] = xs;
          valid |= 1u << (k * VEC + e);
          const K key = KT::key(z[k * VEC + e]);
          kmax = key > kmax ? key : kmax;
        }
      } else {
#pragma unroll
        for (int e = 0; e < VEC; ++e) z[k * VEC + e] = R(0);
      }
    }
    bool hard = false;
    unsigned keepmask = 0;
    if (!select) {
      keepmask = (r >= n) ? valid : 0u;
    } else {
      __syncthreads();
      TR_T(8);  // load
      // ---- largest magnitude of the problem ---------------------------------
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        const K other = __shfl_xor_sync(0xffffffffu, kmax, o);
        kmax = other > kmax ? other : kmax;
      }
      for (int b = t; b < kTrBins; b += kTrThreads) sh->hist[b] = 0;
      if (t == 0) {
        sh->cand_count = 0;
        sh->kmax_slot[crank] = 0;
      }
      __syncthreads();  // also: ws free for reuse
      if (lane == 0) atomicMax(&sh->kmax_slot[crank], (unsigned long long)kmax);
      __syncthreads();
      if (t < (int)csize) cluster.map_shared_rank(sh, t)->kmax_slot[crank] = sh->kmax_slot[crank];
      cluster.sync();
      TR_T(9);  // kmax exchange
      K gmax = 0;
      for (unsigned rr = 0; rr < csize; ++rr) {
        const K v = (K)sh->kmax_slot[rr];
        gmax = v > gmax ? v : gmax;
      }
      // the keys are the bit patterns of |z|: gmax back to a number
      double zmax;
      if (sizeof(R) == 8) zmax = __longlong_as_double((long long)gmax);
      else zmax = (double)__int_as_float((int)gmax);
      const double scale = (double)kTrBins / zmax;
      hard = !(zmax > 0.0) || !(zmax < 1e300) || !(scale < 1e300);  // NaN, Inf, all zero, denormal range
      auto bin_of = [&](R v) -> unsigned {
        const double b = fabs((double)v) * scale;
        const unsigned ub = (unsigned)b;  // b in [0, 2048 (1 + eps)]
        return ub < (unsigned)kTrBins ? ub : (unsigned)kTrBins - 1u;
      };
      long long need = r;
      unsigned sel = 0;
      if (!hard) {
        // ---- one histogram pass ----------------------------------------------
#pragma unroll
        for (int s = 0; s < kTrE; ++s)
          if ((valid >> s) & 1u) atomicAdd(&sh->hist[bin_of(z[s])], 1u);
      }
      __syncthreads();
      TR_T(10);  // histogram
      cluster.sync();  // every CTA's histogram is complete (a hard problem still keeps the barriers aligned)
      if (!hard) {
        for (int b = t; b < kTrBins; b += kTrThreads) {
          unsigned s = 0;
          for (unsigned rr = 0; rr < csize; ++rr) s += cluster.map_shared_rank(sh->hist, rr)[b];
          sh->tot[b] = s;
        }
      }
      cluster.sync();  // remote reads done
      TR_T(11);  // sync + remote sum + sync
      long long bin_count = 0;
      if (!hard) {
        int c[kTrBpt], mine = 0;
#pragma unroll
        for (int j = 0; j < kTrBpt; ++j) {
          c[j] = (int)sh->tot[kTrBins - 1 - kTrBpt * t - j];
          mine += c[j];
        }
        int total;
        long long above = block_excl_scan<kTrThreads>(mine, sh->ws, &total);
#pragma unroll
        for (int j = 0; j < kTrBpt; ++j) {
          if (above < need && need <= above + c[j]) {
            sh->sel_bin = kTrBins - 1 - kTrBpt * t - j;
            sh->sel_above = above;
            sh->sel_count = c[j];
          }
          above += c[j];
        }
        __syncthreads();
        sel = (unsigned)sh->sel_bin;
        need -= sh->sel_above;  // rank inside the threshold bin, 1-based
        bin_count = sh->sel_count;
        hard = bin_count > kTrCandMax;
      }
      TR_T(12);  // scan + pick
      // ---- gather the threshold bin into CTA 0 --------------------------------
      if (!hard) {
#pragma unroll
        for (int k = 0; k < ROUNDS; ++k) {
#pragma unroll
          for (int e = 0; e < VEC; ++e) {
            const int s = k * VEC + e;
            if (((valid >> s) & 1u) && bin_of(z[s]) == sel) {
              const int pos = atomicAdd(&sh0->cand_count, 1);
              sh0->cand_key[pos] = (unsigned long long)KT::key(z[s]);
              sh0->cand_idx[pos] = (int)(cbase + ((long long)k * kTrThreads + t) * VEC + e);
            }
          }
        }
      }
      cluster.sync();
      TR_T(13);  // gather + sync
      if (!hard && crank == 0) {
        // direct ranking: key descending, position ascending; the element of rank `need` is the threshold
        const int C = sh->cand_count;
        for (int ci = t; ci < C; ci += kTrThreads) {
          const unsigned long long mk = sh->cand_key[ci];
          const int mi = sh->cand_idx[ci];
          int rank = 1;
          for (int j = 0; j < C; ++j) {
            const unsigned long long kj = sh->cand_key[j];
            rank += (kj > mk) || (kj == mk && sh->cand_idx[j] < mi);
          }
          if ((long long)rank == need) {
            for (unsigned rr = 0; rr < csize; ++rr) {
              TrShared* dst = cluster.map_shared_rank(sh, rr);
              dst->thr_key = mk;
              dst->thr_idx = mi;
            }
          }
        }
      }
      cluster.sync();
      TR_T(14);  // ranking + sync
      if (!hard) {
        const K tk = (K)sh->thr_key;
        const int ti = sh->thr_idx;
#pragma unroll
        for (int k = 0; k < ROUNDS; ++k) {
#pragma unroll
          for (int e = 0; e < VEC; ++e) {
            const int s = k * VEC + e;
            if ((valid >> s) & 1u) {
              const K key = KT::key(z[s]);
              const int gi = (int)(cbase + ((long long)k * kTrThreads + t) * VEC + e);
              if (key > tk || (key == tk && gi <= ti)) keepmask |= 1u << s;
            }
          }
        }
      }
    }
    if (t == 0 && crank == 0) fallback[prob] = hard ? 1 : 0;

    // ---- write y once (a flagged problem is written by the radix kernel) ------
    if (!hard) {
#pragma unroll
      for (int k = 0; k < ROUNDS; ++k) {
        const long long li = ((long long)k * kTrThreads + t) * VEC;
        if (li < cnt) {
          Pack<R, VEC> o;
#pragma unroll
          for (int e = 0; e < VEC; ++e) {
            const int s = k * VEC + e;
            const R zz = ((keepmask >> s) & 1u) ? z[s] : R(0);
            R v = zz - xs_sm[li + e];  // y .-= xk .+ sj   (:70)
            if (BINF) v = jl_min(jl_max(v, -delta), delta);  // shiftedIndBallL0BInf.jl:91
            o.v[e] = v;
          }
          st_stream(y + pbase + cbase + li, o);
        }
      }
    }
    cluster.sync();  // shared state is reused by the next problem
    TR_T(15);  // keep masks + write + sync
  }
}

// ------------------------------------------------- global multi-pass path ---
// For one long vector: z is stashed in y (3R + 1W), every further digit costs
// one read of y, the last pass reads y, xk, sj and writes y.
struct GlobalSel {
  unsigned long long hist[kTrBins];  // 64-bit: the sharded form sums the histograms of every GPU
  unsigned long long prefix;
  long long need;
  int shift;
  int done;          // keep_all_bin reached
  long long eq_total;
  long long eq_base;  // threshold-equal elements on lower-ranked shards (sharded form)
};

template <class R>
__global__ void __launch_bounds__(256) topr_g_stash(R* y, const R* xk, const R* sj, const R* q, long long n,
                                                    GlobalSel* st) {
  using KT = KeyTraits<R>;
  __shared__ unsigned h[kTrBins];
  for (int b = threadIdx.x; b < kTrBins; b += 256) h[b] = 0;
  __syncthreads();
  const int width = digit_bits(KT::BITS, 0);
  const int shift = KT::BITS - width;
  for (long long i = (long long)blockIdx.x * 256 + threadIdx.x; i < n; i += (long long)gridDim.x * 256) {
    const R z = (xk[i] + sj[i]) + q[i];
    y[i] = z;
    atomicAdd(&h[(unsigned)(KT::key(z) >> shift)], 1u);
  }
  __syncthreads();
  for (int b = threadIdx.x; b < kTrBins; b += 256)
    if (h[b]) atomicAdd(&st->hist[b], (unsigned long long)h[b]);
}

template <class R>
__global__ void __launch_bounds__(256) topr_g_hist(const R* y, long long n, int pass, GlobalSel* st) {
  using KT = KeyTraits<R>;
  using K = typename KT::K;
  __shared__ unsigned h[kTrBins];
  if (st->done) return;
  for (int b = threadIdx.x; b < kTrBins; b += 256) h[b] = 0;
  __syncthreads();
  const int width = digit_bits(KT::BITS, pass);
  const int hi_shift = st->shift;
  const int shift = hi_shift - width;
  const K prefix = (K)st->prefix;
  for (long long i = (long long)blockIdx.x * 256 + threadIdx.x; i < n; i += (long long)gridDim.x * 256) {
    const K key = KT::key(y[i]);
    if ((key >> hi_shift) == prefix) atomicAdd(&h[(unsigned)((key >> shift) & (K)((1u << width) - 1u))], 1u);
  }
  __syncthreads();
  for (int b = threadIdx.x; b < kTrBins; b += 256)
    if (h[b]) atomicAdd(&st->hist[b], (unsigned long long)h[b]);
}

// single block: pick the digit, update the selection state, clear the histogram
template <class R>
__global__ void __launch_bounds__(kPickThreads) topr_g_pick(int pass, GlobalSel* st) {
  using KT = KeyTraits<R>;
  __shared__ int ws[40];
  __shared__ int sel_bin;
  __shared__ long long sel_above, sel_count;
  if (st->done) return;
  const int t = threadIdx.x;
  const int width = digit_bits(KT::BITS, pass);
  const int b0 = kTrBins - 1 - 2 * t, b1 = b0 - 1;
  const long long c0 = (long long)st->hist[b0], c1 = (long long)st->hist[b1];
  // counts can exceed int for huge n: scan in two 31-bit halves is overkill; clamp-free 64-bit scan
  // via two int scans of the low / high parts
  int total_lo, total_hi;
  const long long s = c0 + c1;
  const int lo = (int)(s & 0x3fffffff), hi = (int)(s >> 30);
  const long long above_lo = block_excl_scan<kPickThreads>(lo, ws, &total_lo);
  const long long above_hi = block_excl_scan<kPickThreads>(hi, ws, &total_hi);
  const long long above = above_lo + (above_hi << 30);
  const long long need = st->need;
  if (above < need && need <= above + c0) {
    sel_bin = b0; sel_above = above; sel_count = c0;
  } else if (above + c0 < need && need <= above + c0 + c1) {
    sel_bin = b1; sel_above = above + c0; sel_count = c1;
  }
  __syncthreads();
  st->hist[b0] = 0;
  st->hist[b1] = 0;
  if (t == 0) {
    st->prefix = (st->prefix << width) | (unsigned long long)sel_bin;
    st->need = need - sel_above;
    st->shift -= width;
    if (st->need == sel_count) st->done = 1;
  }
}

// per-block count of threshold-equal elements (block-contiguous slices, index order)
template <class R>
__global__ void __launch_bounds__(256) topr_g_eqcount(const R* y, long long n, long long per_block,
                                                      const GlobalSel* st, long long* block_eq) {
  using KT = KeyTraits<R>;
  using K = typename KT::K;
  if (st->done) {
    if (threadIdx.x == 0) block_eq[blockIdx.x] = 0;
    return;
  }
  const long long b = (long long)blockIdx.x * per_block;
  const long long e = b + per_block < n ? b + per_block : n;
  const K prefix = (K)st->prefix;
  long long c = 0;
  for (long long i = b + threadIdx.x; i < e; i += 256) c += (KT::key(y[i]) >> st->shift) == prefix;
  __shared__ long long shc[8];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
  if ((threadIdx.x & 31) == 0) shc[threadIdx.x >> 5] = c;
  __syncthreads();
  if (threadIdx.x == 0) {
    long long s = 0;
    for (int w = 0; w < 8; ++w) s += shc[w];
    block_eq[blockIdx.x] = s;
  }
}
__global__ void topr_g_scan(long long* block_eq, int nblocks, GlobalSel* st) {
  if (threadIdx.x == 0 && blockIdx.x == 0) {
    long long run = 0;
    for (int i = 0; i < nblocks; ++i) {
      long long c = block_eq[i];
      block_eq[i] = run;
      run += c;
    }
    st->eq_total = run;
  }
}

template <class R, bool BINF>
__global__ void __launch_bounds__(256) topr_g_final(R* y, const R* xk, const R* sj, long long n, long long per_block,
                                                    const GlobalSel* st, const long long* block_eq, R delta,
                                                    int mode /*0 select, 1 keep all, 2 keep none*/) {
  using KT = KeyTraits<R>;
  using K = typename KT::K;
  __shared__ int wsum[8];
  const long long b = (long long)blockIdx.x * per_block;
  const long long e = b + per_block < n ? b + per_block : n;
  const K prefix = (K)st->prefix;
  const int shift = st->shift;
  const bool all_bin = st->done != 0;
  const long long need = st->need;
  long long run = mode == 0 ? st->eq_base + block_eq[blockIdx.x] : 0;
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  for (long long base = b; base < e; base += 256) {
    const long long i = base + threadIdx.x;
    R z = R(0), xs = R(0);
    bool eq = false, keep = false;
    if (i < e) {
      z = y[i];
      xs = xk[i] + sj[i];
      if (mode == 1) keep = true;
      else if (mode == 0) {
        const K kp = KT::key(z) >> shift;
        keep = kp > prefix;
        eq = kp == prefix;
      }
    }
    if (mode == 0) {
      if (all_bin) {
        keep = keep || eq;
      } else {
        const unsigned bal = __ballot_sync(0xffffffffu, eq);
        if (lane == 0) wsum[w] = __popc(bal);
        __syncthreads();
        long long off = run;
        int tot = 0;
        for (int ww = 0; ww < 8; ++ww) {
          if (ww < w) off += wsum[ww];
          tot += wsum[ww];
        }
        const long long rank = off + __popc(bal & ((1u << lane) - 1u));
        if (eq && rank < need) keep = true;
        run += tot;
        __syncthreads();
      }
    }
    if (i < e) {
      R v = (keep ? z : R(0)) - xs;
      if (BINF) v = jl_min(jl_max(v, -delta), delta);
      y[i] = v;
    }
  }
}

// One vector on one GPU (reduce == nullptr), or this GPU's contiguous shard of a vector spread over
// `world` GPUs in rank order: the histogram of every digit is summed over the shards through the caller's
// all-reduce (one exchange per radix digit), every rank picks the same digit, and the
// lowest-index tie rule continues across shards through the count of threshold-equal elements on the
// lower-ranked shards.
template <class R>
static int32_t topr_global(spx_ctx* ctx, int64_t n, R* y, const R* xk, const R* sj, const R* q, int64_t r, bool binf,
                           R delta, int64_t n_global = -1, int rank = 0, int world = 1,
                           spx_allreduce_sum_fn reduce = nullptr, void* user = nullptr) {
  using KT = KeyTraits<R>;
  if (n_global < 0) n_global = n;
  const int nblk = ctx->sm_count * 8;
  int32_t stt = ensure_scratch(ctx, sizeof(GlobalSel) + sizeof(long long) * (size_t)(nblk + 1));
  if (stt != SPX_OK) return stt;
  GlobalSel* st = (GlobalSel*)ctx->d_scratch;
  long long* block_eq = (long long*)((char*)ctx->d_scratch + sizeof(GlobalSel));
  GlobalSel init;
  memset(&init, 0, sizeof(init));
  init.need = r;
  init.shift = KT::BITS;
  SPX_CUDA(cudaMemcpyAsync(st, &init, sizeof(init), cudaMemcpyHostToDevice, ctx->stream));
  const int mode = (r >= n_global) ? 1 : (r <= 0 ? 2 : 0);
  // sum the histogram over the shards (counts are exact in Float64 up to 2^53)
  std::vector<unsigned long long> hh;
  std::vector<double> hd;
  auto reduce_hist = [&]() -> int32_t {
    if (!reduce) return SPX_OK;
    hh.resize(kTrBins);
    hd.resize(kTrBins);
    SPX_CUDA(cudaMemcpyAsync(hh.data(), st->hist, sizeof(unsigned long long) * kTrBins, cudaMemcpyDeviceToHost,
                             ctx->stream));
    SPX_CUDA(cudaStreamSynchronize(ctx->stream));
    for (int b = 0; b < kTrBins; ++b) hd[b] = (double)hh[b];
    int32_t st2 = reduce(user, hd.data(), kTrBins);
    if (st2 != SPX_OK) {
      set_error("spx_prox_indballl0_sharded: all-reduce callback failed (%d)", (int)st2);
      return st2;
    }
    for (int b = 0; b < kTrBins; ++b) hh[b] = (unsigned long long)hd[b];
    SPX_CUDA(cudaMemcpyAsync(st->hist, hh.data(), sizeof(unsigned long long) * kTrBins, cudaMemcpyHostToDevice,
                             ctx->stream));
    return SPX_OK;
  };
  if (n > 0) {
    topr_g_stash<R><<<nblk, 256, 0, ctx->stream>>>(y, xk, sj, q, n, st);
    ctx->launches++;
  }
  const long long per_block = (n + nblk - 1) / nblk;
  if (mode == 0) {
    for (int pass = 0; pass < KT::NPASS; ++pass) {
      if (pass > 0 && n > 0) {
        topr_g_hist<R><<<nblk, 256, 0, ctx->stream>>>(y, n, pass, st);
        ctx->launches++;
      }
      stt = reduce_hist();
      if (stt != SPX_OK) return stt;
      topr_g_pick<R><<<1, kPickThreads, 0, ctx->stream>>>(pass, st);
      ctx->launches++;
    }
    topr_g_eqcount<R><<<nblk, 256, 0, ctx->stream>>>(y, n, per_block, st, block_eq);
    topr_g_scan<<<1, 32, 0, ctx->stream>>>(block_eq, nblk, st);
    ctx->launches += 2;
    if (reduce && world > 1) {
      // exclusive prefix over ranks of the threshold-equal counts: all-reduce a vector with one slot per rank
      long long eq_total = 0;
      SPX_CUDA(cudaMemcpyAsync(&eq_total, &st->eq_total, sizeof(long long), cudaMemcpyDeviceToHost, ctx->stream));
      SPX_CUDA(cudaStreamSynchronize(ctx->stream));
      std::vector<double> slots((size_t)world, 0.0);
      slots[(size_t)rank] = (double)eq_total;
      int32_t st2 = reduce(user, slots.data(), world);
      if (st2 != SPX_OK) {
        set_error("spx_prox_indballl0_sharded: all-reduce callback failed (%d)", (int)st2);
        return st2;
      }
      long long base = 0;
      for (int rr = 0; rr < rank; ++rr) base += (long long)slots[(size_t)rr];
      SPX_CUDA(cudaMemcpyAsync(&st->eq_base, &base, sizeof(long long), cudaMemcpyHostToDevice, ctx->stream));
      SPX_CUDA(cudaStreamSynchronize(ctx->stream));  // `base` lives on this stack frame
    }
  }
  if (n > 0) {
    if (binf)
      topr_g_final<R, true><<<nblk, 256, 0, ctx->stream>>>(y, xk, sj, n, per_block, st, block_eq, delta, mode);
    else
      topr_g_final<R, false><<<nblk, 256, 0, ctx->stream>>>(y, xk, sj, n, per_block, st, block_eq, delta, mode);
    ctx->launches++;
  }
  SPX_CUDA(cudaGetLastError());
  return SPX_OK;
}

template <class Kern>
static int32_t cluster_grid(spx_ctx* ctx, Kern kern, int csize, size_t smem, int64_t nprob, cudaLaunchConfig_t* cfg,
                            cudaLaunchAttribute* attr) {
  SPX_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = csize;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg->blockDim = dim3(kTrThreads);
  cfg->dynamicSmemBytes = smem;
  cfg->stream = ctx->stream;
  cfg->attrs = attr;
  cfg->numAttrs = 1;
  cfg->gridDim = dim3(csize);
  int max_clusters = 0;
  cudaError_t e = cudaOccupancyMaxActiveClusters(&max_clusters, kern, cfg);
  if (e != cudaSuccess || max_clusters < 1) max_clusters = ctx->sm_count / csize;
  if (max_clusters < 1) max_clusters = 1;
  long long nclusters = nprob < max_clusters ? nprob : max_clusters;
  cfg->gridDim = dim3((unsigned)(nclusters * csize));
  return SPX_OK;
}

template <class R, bool BINF, bool VECLD>
static int32_t topr_cluster_launch(spx_ctx* ctx, int64_t nprob, int64_t n, R* y, const R* xk, const R* sj, const R* q,
                                   int64_t r, R delta) {
  int csize = 1;
  while ((long long)csize * kTrChunk < n) csize <<= 1;
  const size_t smem = ((sizeof(TrShared) + 15) / 16) * 16 + sizeof(R) * (size_t)kTrChunk;
  int32_t st = ensure_scratch(ctx, (size_t)nprob + 4096);
  if (st != SPX_OK) return st;
  unsigned char* flags = (unsigned char*)ctx->d_scratch + 4096;  // the first 4 KiB belong to the reductions
  cudaLaunchConfig_t cfg = {};
  cudaLaunchAttribute attr[1];
  {  // linear-bin select: decides every ordinary problem
    auto kern = topr_cluster_lin_kernel<R, BINF, VECLD>;
    st = cluster_grid(ctx, kern, csize, smem, nprob, &cfg, attr);
    if (st != SPX_OK) return st;
    SPX_CUDA(cudaLaunchKernelEx(&cfg, kern, y, xk, sj, q, (long long)n, (long long)r, delta, (long long)nprob, flags));
    ctx->launches++;
  }
  {  // radix select on the problems it flagged (NaN / Inf / all-zero / crowded threshold bin)
    auto kern = topr_cluster_kernel<R, BINF, VECLD>;
    st = cluster_grid(ctx, kern, csize, smem, nprob, &cfg, attr);
    if (st != SPX_OK) return st;
    SPX_CUDA(cudaLaunchKernelEx(&cfg, kern, y, xk, sj, q, (long long)n, (long long)r, delta, (long long)nprob,
                                (const unsigned char*)flags));
    ctx->launches++;
  }
  return SPX_OK;
}

template <class R>
static int32_t prox_indballl0(spx_ctx* ctx, int64_t nprob, int64_t n, R* y, const R* xk, const R* sj, const R* q,
                              int64_t r, int32_t binf, double delta) {
  SPX_REQUIRE(ctx != nullptr, "null context");
  SPX_REQUIRE(nprob >= 0 && n >= 0, "negative size");
  if (nprob == 0 || n == 0) return SPX_OK;
  SPX_REQUIRE(y && xk && sj && q, "null device vector");
  DeviceGuard g(ctx->device);
  if (n > (int64_t)kTrMaxCluster * kTrChunk) {
    for (int64_t p = 0; p < nprob; ++p) {
      int32_t st = topr_global<R>(ctx, n, y + p * n, xk + p * n, sj + p * n, q + p * n, r, binf != 0, (R)delta);
      if (st != SPX_OK) return st;
    }
    return SPX_OK;
  }
  const uintptr_t bits = (uintptr_t)y | (uintptr_t)xk | (uintptr_t)sj | (uintptr_t)q;
  const bool vec = (bits & 15u) == 0 && (n * (int64_t)sizeof(R)) % 16 == 0;
  if (binf) {
    return vec ? topr_cluster_launch<R, true, true>(ctx, nprob, n, y, xk, sj, q, r, (R)delta)
               : topr_cluster_launch<R, true, false>(ctx, nprob, n, y, xk, sj, q, r, (R)delta);
  }
  return vec ? topr_cluster_launch<R, false, true>(ctx, nprob, n, y, xk, sj, q, r, (R)delta)
             : topr_cluster_launch<R, false, false>(ctx, nprob, n, y, xk, sj, q, r, (R)delta);
}

}  // namespace spx

using namespace spx;

#ifdef SPX_TR_TIMING
extern "C" int32_t spx_debug_topr_timing(unsigned long long* out8, int reset) {
  cudaMemcpyFromSymbol(out8, g_tr_t, 128);
  if (reset) {
    unsigned long long z[16] = {0};
    cudaMemcpyToSymbol(g_tr_t, z, 128);
  }
  return 0;
}
#endif
extern "C" int32_t spx_prox_indballl0_f64(spx_ctx* ctx, int64_t nprob, int64_t n, double* y, const double* xk,
                                          const double* sj, const double* q, int64_t r, int32_t binf, double delta) {
  return prox_indballl0<double>(ctx, nprob, n, y, xk, sj, q, r, binf, delta);
}
extern "C" int32_t spx_prox_indballl0_sharded_f64(spx_ctx* ctx, int64_t n_local, int64_t n_global, double* y,
                                                  const double* xk, const double* sj, const double* q, int64_t r,
                                                  int32_t binf, double delta, int32_t rank, int32_t world,
                                                  spx_allreduce_sum_fn reduce, void* user) {
  SPX_REQUIRE(ctx != nullptr, "null context");
  SPX_REQUIRE(n_local >= 0 && n_global >= n_local && world >= 1 && rank >= 0 && rank < world, "bad shard description");
  SPX_REQUIRE(n_local == 0 || (y && xk && sj && q), "null device vector");
  SPX_REQUIRE(world == 1 || reduce != nullptr, "a sharded vector needs the all-reduce callback");
  DeviceGuard g(ctx->device);
  return topr_global<double>(ctx, n_local, y, xk, sj, q, r, binf != 0, delta, n_global, rank, world, reduce, user);
}
extern "C" int32_t spx_prox_indballl0_sharded_f32(spx_ctx* ctx, int64_t n_local, int64_t n_global, float* y,
                                                  const float* xk, const float* sj, const float* q, int64_t r,
                                                  int32_t binf, double delta, int32_t rank, int32_t world,
                                                  spx_allreduce_sum_fn reduce, void* user) {
  SPX_REQUIRE(ctx != nullptr, "null context");
  SPX_REQUIRE(n_local >= 0 && n_global >= n_local && world >= 1 && rank >= 0 && rank < world, "bad shard description");
  SPX_REQUIRE(n_local == 0 || (y && xk && sj && q), "null device vector");
  SPX_REQUIRE(world == 1 || reduce != nullptr, "a sharded vector needs the all-reduce callback");
  DeviceGuard g(ctx->device);
  return topr_global<float>(ctx, n_local, y, xk, sj, q, r, binf != 0, (float)delta, n_global, rank, world, reduce,
                            user);
}
extern "C" int32_t spx_prox_indballl0_f32(spx_ctx* ctx, int64_t nprob, int64_t n, float* y, const float* xk,
                                          const float* sj, const float* q, int64_t r, int32_t binf, double delta) {
  return prox_indballl0<float>(ctx, nprob, n, y, xk, sj, q, r, binf, delta);
}
