// hybrid_fuse.cu — dense + BM25 score fusion for the hybrid top-k (BASELINE config 4 / the metric).
//
// The reference has NO dense+BM25 fusion formula (its "HybridRetriever" is vector search + LLM
// rerank, src/retrieval.py:297-429; BM25Retriever is never instantiated).  This kernel applies the
// reference's own aggregation rule (src/retrieval.py:1639-1661) with BM25 as one more
// retrieval source — PARITY UNPINNED, checked against oracle/fusion.py (our restatement):
//   candidates = dense top-k  U  BM25 top-k                      (keyed by chunk id)
//   base       = dense inner product if the chunk is a dense hit, else the BM25 score min-max
//                normalised over this query's BM25 hits ((s - min)/(max - min); a constant list
//                maps to 1 if max > 0 else 0)
//   final      = base * (1 + 0.1 * (n_methods - 1))              (method-diversity bonus, :1653-1657;
//                                                                 one query text => hit-count bonus 1)
//   result     = k best by (final descending, id ascending)
// All arithmetic in float64, result stored as float32.  One CTA per query; <= 2 * 1024 candidates.
#include <algorithm>

#include "b2r_common.cuh"

namespace b2r {

// Fusion of one query's dense hits (Dq, Iq) and BM25 hits (Sq, Jq); whole-CTA routine.  The inputs
// may live in global or shared memory.  buf: shared scratch of buf_cap keys (power of two >=
// k_dense + k_bm25 and >= k_out); red: shared double[2].
__device__ __forceinline__ void fuse_query(const float* Dq, const long long* Iq, const double* Sq,
                                           const long long* Jq, int k_dense, int k_bm25, int k_out, float* F,
                                           long long* O, unsigned char* M, u64* buf, int buf_cap, double* red) {
  const int tid = threadIdx.x, nt = blockDim.x;
  if (tid == 0) {
    double mn = 1.7976931348623157e308, mx = -1.7976931348623157e308;
    for (int j = 0; j < k_bm25; ++j) {
      if (Jq[j] >= 0) {
        mn = Sq[j] < mn ? Sq[j] : mn;
        mx = Sq[j] > mx ? Sq[j] : mx;
      }
    }
    red[0] = mn;
    red[1] = mx;
  }
  for (int i = tid; i < buf_cap; i += nt) buf[i] = 0;
  __syncthreads();
  const double mn = red[0], mx = red[1];
  for (int i = tid; i < k_dense + k_bm25; i += nt) {
    u64 key = 0;
    if (i < k_dense) {
      const long long id = Iq[i];
      if (id >= 0) {
        bool both = false;
        for (int j = 0; j < k_bm25; ++j) both |= (Jq[j] == id);
        const double fin = __dmul_rn(static_cast<double>(Dq[i]), both ? 1.1 : 1.0);
        key = make_key(static_cast<float>(fin), static_cast<u32>(id));
      }
    } else {
      const int j = i - k_dense;
      const long long id = Jq[j];
      if (id >= 0) {
        bool dense_too = false;
        for (int t = 0; t < k_dense; ++t) dense_too |= (Iq[t] == id);
        if (!dense_too) {
          const double norm = (mx > mn) ? __ddiv_rn(__dsub_rn(Sq[j], mn), __dsub_rn(mx, mn)) : (mx > 0.0 ? 1.0 : 0.0);
          key = make_key(static_cast<float>(norm), static_cast<u32>(id));
        }
      }
    }
    buf[i] = key;
  }
  __syncthreads();
  block_sort_desc(buf, buf_cap);
  for (int i = tid; i < k_out; i += nt) {
    const u64 key = buf[i];
    if (key == 0) {
      F[i] = -3.4028235e38f;
      O[i] = -1;
      if (M) M[i] = 0;
    } else {
      const long long id = static_cast<long long>(key_id(key));
      F[i] = key_score(key);
      O[i] = id;
      if (M) {
        unsigned char m = 0;
        for (int t = 0; t < k_dense; ++t) m |= (Iq[t] == id) ? 1 : 0;
        for (int j = 0; j < k_bm25; ++j) m |= (Jq[j] == id) ? 2 : 0;
        M[i] = m;  // bit 0: dense hit, bit 1: BM25 hit  (-> retrieval_sources of the record)
      }
    }
  }
}

__global__ void __launch_bounds__(256) hybrid_fuse_kernel(const float* __restrict__ D, const long long* __restrict__ I,
                                                          const double* __restrict__ S, const long long* __restrict__ J,
                                                          int k_dense, int k_bm25, int k_out, float* __restrict__ F,
                                                          long long* __restrict__ O, unsigned char* __restrict__ M,
                                                          int buf_cap) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  // a scan launched behind this kernel with B2R_LAUNCH_OVERLAP_PREVIOUS (the next hybrid step) may start its pass
  // over the rows now; it writes nothing before this kernel has completed
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  u64* buf = reinterpret_cast<u64*>(smem_raw);             // [buf_cap] keys
  double* red = reinterpret_cast<double*>(buf + buf_cap);  // [2] min / max
  const size_t q = blockIdx.x;
  fuse_query(D + q * k_dense, I + q * k_dense, S + q * k_bm25, J + q * k_bm25, k_dense, k_bm25, k_out, F + q * k_out,
             O + q * k_out, M ? M + q * k_out : nullptr, buf, buf_cap, red);
}

// ----------------------------------------------------------------------------------------------
// Multi-GPU: exchange + merge + fusion in ONE kernel over NVLink peer memory.
//
// Every rank owns a symmetric buffer (same layout on all ranks, mapped into every peer):
//   slots[2 parities][world sources][ D f32[nq_max*k_max] | I i64[..] | S f64[..] | J i64[..] ]
//   flags[2 parities][world sources][nq_max] u32
// CTA q of rank r: (1) stores its rank's hits for query q into slot [parity][r] of EVERY rank's
// buffer (plain stores to peer-mapped addresses travel over NVLink), (2) system-scope fence, then
// one thread publishes flag[parity][r][q] = step on every rank (st.release.sys), (3) waits until
// its own flags [parity][*][q] all equal `step` (ld.acquire.sys), (4) merges the world lists
// (list-major tie order = ascending global id for row shards) and fuses dense with BM25.
// It replaces all-gather + 2 merge launches + the fusion launch of the NCCL path; every rank ends
// with the same answer.  Parity = step & 1 double-buffers the slots: a rank can only reach step
// s + 2 after all peers published step s + 1, i.e. after they finished reading step s.
// ----------------------------------------------------------------------------------------------
struct ExchangeArgs {
  unsigned char* peers[8];
  int rank, world;
  int nq_max, k_max;
  u32 step;
  u32* status;              // device-visible word (may be mapped host memory), 0 = ok; NULL: a timed-out wait traps
  unsigned long long timeout_ns;  // bound on the wait for the peers' flags
};

__host__ __device__ inline size_t exch_slot_bytes(int nq_max, int k_max) {
  return static_cast<size_t>(nq_max) * k_max * 28;
}
__host__ __device__ inline size_t exch_flags_offset(int nq_max, int k_max, int world) {
  return 2 * static_cast<size_t>(world) * exch_slot_bytes(nq_max, k_max);
}
__host__ __device__ inline size_t exch_total_bytes(int nq_max, int k_max, int world) {
  return exch_flags_offset(nq_max, k_max, world) + 2 * static_cast<size_t>(world) * nq_max * sizeof(u32);
}

__device__ __forceinline__ void st_release_sys(u32* p, u32 v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ u32 ld_acquire_sys(const u32* p) {
  u32 v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ unsigned long long global_timer_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}

__global__ void __launch_bounds__(256) hybrid_exchange_fuse_kernel(const float* __restrict__ D,
                                                                   const long long* __restrict__ I,
                                                                   const double* __restrict__ S,
                                                                   const long long* __restrict__ J, int nq, int k,
                                                                   const ExchangeArgs x, float* __restrict__ F,
                                                                   long long* __restrict__ O) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  __shared__ int timed_out;
  // the next step's scan (B2R_LAUNCH_OVERLAP_PREVIOUS) may start now: this kernel's wait for the slowest rank then
  // costs nothing, the fast ranks are already streaming their next query
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  const int q = blockIdx.x, tid = threadIdx.x, nt = blockDim.x;
  if (tid == 0) timed_out = 0;
  const int parity = static_cast<int>(x.step & 1u);
  const size_t slot = exch_slot_bytes(x.nq_max, x.k_max);
  const size_t nk = static_cast<size_t>(x.nq_max) * x.k_max;
  const size_t flags_off = exch_flags_offset(x.nq_max, x.k_max, x.world);
  const size_t qoff = static_cast<size_t>(q) * x.k_max;

  // (1) push my rank's lists for query q to every rank (my own buffer included)
  for (int p = 0; p < x.world; ++p) {
    unsigned char* base = x.peers[p] + (static_cast<size_t>(parity) * x.world + x.rank) * slot;
    float* pD = reinterpret_cast<float*>(base) + qoff;
    long long* pI = reinterpret_cast<long long*>(base + nk * 4) + qoff;
    double* pS = reinterpret_cast<double*>(base + nk * 12) + qoff;
    long long* pJ = reinterpret_cast<long long*>(base + nk * 20) + qoff;
    for (int i = tid; i < k; i += nt) {
      pD[i] = D[static_cast<size_t>(q) * k + i];
      pI[i] = I[static_cast<size_t>(q) * k + i];
      pS[i] = S[static_cast<size_t>(q) * k + i];
      pJ[i] = J[static_cast<size_t>(q) * k + i];
    }
  }
  __threadfence_system();
  __syncthreads();
  // (2) publish, (3) wait for everybody's lists
  if (tid < x.world) {
    u32* pf = reinterpret_cast<u32*>(x.peers[tid] + flags_off) + (static_cast<size_t>(parity) * x.world + x.rank) * x.nq_max + q;
    st_release_sys(pf, x.step);
    const u32* mine = reinterpret_cast<const u32*>(x.peers[x.rank] + flags_off) +
                      (static_cast<size_t>(parity) * x.world + tid) * x.nq_max + q;
    // bounded wait: a peer that failed before launching (or runs a different step) must not hang this GPU forever
    const unsigned long long t0 = global_timer_ns();
    u32 spins = 0;
    while (ld_acquire_sys(mine) != x.step) {
      if ((++spins & 1023u) == 0 && global_timer_ns() - t0 > x.timeout_ns) {
        if (x.status == nullptr) __trap();
        *reinterpret_cast<volatile u32*>(x.status) = 1u + static_cast<u32>(tid);   // 1 + a rank found missing
        timed_out = 1;
        break;
      }
    }
  }
  __syncthreads();
  if (timed_out) {  // no answer for this query: pad, and let the host report B2R_ERR_CUDA from the status word
    for (int i = tid; i < k; i += nt) {
      F[static_cast<size_t>(q) * k + i] = -3.4028235e38f;
      O[static_cast<size_t>(q) * k + i] = -1;
    }
    return;
  }

  // (4) merge the world lists of this query, then fuse
  const int total = x.world * k;
  int cap = 2;
  while (cap < total) cap <<= 1;
  int fcap = 2;
  while (fcap < 2 * k) fcap <<= 1;
  u128* kb = reinterpret_cast<u128*>(smem_raw);                         // [cap] BM25 keys (also reused for dense keys)
  float* mD = reinterpret_cast<float*>(kb + cap);                       // merged dense scores [k]
  long long* mI = reinterpret_cast<long long*>(mD + ((k + 1) & ~1));    // [k]
  double* mS = reinterpret_cast<double*>(mI + k);                       // [k]
  long long* mJ = reinterpret_cast<long long*>(mS + k);                 // [k]
  double* red = reinterpret_cast<double*>(mJ + k);                      // [2]
  u64* fbuf = reinterpret_cast<u64*>(red + 2);                          // [fcap]
  const unsigned char* my = x.peers[x.rank] + static_cast<size_t>(parity) * x.world * slot;
  // dense: 64-bit keys (score, position); position order == (rank, slot) == ascending global id on ties
  u64* kd = reinterpret_cast<u64*>(kb);
  for (int g = tid; g < cap; g += nt) {
    u64 key = 0;
    if (g < total) {
      const int l = g / k, s = g - l * k;
      const unsigned char* b = my + static_cast<size_t>(l) * slot;
      const long long id = __ldcg(reinterpret_cast<const long long*>(b + nk * 4) + qoff + s);
      if (id >= 0) key = make_key(__ldcg(reinterpret_cast<const float*>(b) + qoff + s), static_cast<u32>(g));
    }
    kd[g] = key;
  }
  __syncthreads();
  block_sort_desc(kd, cap);
  for (int i = tid; i < k; i += nt) {
    const u64 key = kd[i];
    if (key == 0) {
      mD[i] = -3.4028235e38f;
      mI[i] = -1;
    } else {
      const int g = static_cast<int>(key_id(key)), l = g / k, s = g - l * k;
      const unsigned char* b = my + static_cast<size_t>(l) * slot;
      mD[i] = __ldcg(reinterpret_cast<const float*>(b) + qoff + s);
      mI[i] = __ldcg(reinterpret_cast<const long long*>(b + nk * 4) + qoff + s);
    }
  }
  __syncthreads();
  // BM25: 128-bit keys (float64 score, position)
  for (int g = tid; g < cap; g += nt) {
    u128 key = 0;
    if (g < total) {
      const int l = g / k, s = g - l * k;
      const unsigned char* b = my + static_cast<size_t>(l) * slot;
      const long long id = __ldcg(reinterpret_cast<const long long*>(b + nk * 20) + qoff + s);
      if (id >= 0) {
        double sc = __ldcg(reinterpret_cast<const double*>(b + nk * 12) + qoff + s) + 0.0;
        const u64 u = static_cast<u64>(__double_as_longlong(sc));
        const u64 o = (u >> 63) ? ~u : (u | 0x8000000000000000ull);
        key = (static_cast<u128>(o) << 64) | static_cast<u128>(0xFFFFFFFFu - static_cast<u32>(g));
      }
    }
    kb[g] = key;
  }
  __syncthreads();
  block_sort_desc(kb, cap);
  for (int i = tid; i < k; i += nt) {
    const u128 key = kb[i];
    if (key == 0) {
      mS[i] = -1.7976931348623157e308;
      mJ[i] = -1;
    } else {
      const int g = static_cast<int>(0xFFFFFFFFu - static_cast<u32>(key)), l = g / k, s = g - l * k;
      const unsigned char* b = my + static_cast<size_t>(l) * slot;
      mS[i] = __ldcg(reinterpret_cast<const double*>(b + nk * 12) + qoff + s);
      mJ[i] = __ldcg(reinterpret_cast<const long long*>(b + nk * 20) + qoff + s);
    }
  }
  __syncthreads();
  fuse_query(mD, mI, mS, mJ, k, k, k, F + static_cast<size_t>(q) * k, O + static_cast<size_t>(q) * k, nullptr, fbuf, fcap,
             red);
}

}  // namespace b2r

using namespace b2r;

extern "C" int b2r_hybrid_fuse_dev(const float* D, const int64_t* I, const double* S, const int64_t* J, int nq,
                                   int k_dense, int k_bm25, int k_out, float* F, int64_t* O, unsigned char* methods,
                                   void* stream) {
  if (nq < 0 || k_dense < 0 || k_bm25 < 0 || k_out < 1 || k_dense + k_bm25 < 1) return fail(B2R_ERR_INVALID, "b2r_hybrid_fuse: bad sizes");
  if (k_dense > B2R_MAX_K || k_bm25 > B2R_MAX_K || k_out > B2R_MAX_K) return fail(B2R_ERR_UNSUPPORTED, "b2r_hybrid_fuse: k > B2R_MAX_K (1024)");
  if (nq == 0) return B2R_OK;
  if (!F || !O || (k_dense && (!D || !I)) || (k_bm25 && (!S || !J))) return fail(B2R_ERR_INVALID, "b2r_hybrid_fuse: NULL buffer");
  int cap = 2;
  while (cap < std::max(k_dense + k_bm25, k_out)) cap <<= 1;
  const size_t smem = static_cast<size_t>(cap) * 8 + 16;
  hybrid_fuse_kernel<<<nq, 256, smem, static_cast<cudaStream_t>(stream)>>>(
      D, reinterpret_cast<const long long*>(I), S, reinterpret_cast<const long long*>(J), k_dense, k_bm25, k_out, F,
      reinterpret_cast<long long*>(O), methods, cap);
  B2R_CUDA(cudaGetLastError());
  return B2R_OK;
}

extern "C" size_t b2r_hybrid_exchange_bytes(int nq_max, int k_max, int world) {
  if (nq_max < 1 || k_max < 1 || world < 1) return 0;
  return exch_total_bytes(nq_max, k_max, world);
}

extern "C" int b2r_hybrid_exchange_fuse_dev(const float* D, const int64_t* I, const double* S, const int64_t* J, int nq,
                                            int k, void* const* peer_bufs, int rank, int world, int nq_max, int k_max,
                                            uint32_t step, float* F, int64_t* O, uint32_t* status, void* stream) {
  if (nq < 0 || k < 1 || world < 1 || world > 8 || rank < 0 || rank >= world) return fail(B2R_ERR_INVALID, "b2r_hybrid_exchange: bad sizes");
  if (nq > nq_max || k > k_max || k > 256) return fail(B2R_ERR_INVALID, "b2r_hybrid_exchange: nq / k exceed the buffer geometry");
  if (step == 0) return fail(B2R_ERR_INVALID, "b2r_hybrid_exchange: step must be >= 1 (0 is the cleared flag value)");
  if (nq == 0) return B2R_OK;
  if (!D || !I || !S || !J || !F || !O || !peer_bufs) return fail(B2R_ERR_INVALID, "b2r_hybrid_exchange: NULL buffer");
  ExchangeArgs x;
  for (int p = 0; p < 8; ++p) x.peers[p] = p < world ? static_cast<unsigned char*>(peer_bufs[p]) : nullptr;
  x.rank = rank;
  x.world = world;
  x.nq_max = nq_max;
  x.k_max = k_max;
  x.step = step;
  x.status = status;
  const int tmo_ms = tunable("exchange.timeout_ms");
  x.timeout_ns = static_cast<unsigned long long>(tmo_ms > 0 ? tmo_ms : 10000) * 1000000ull;
  int cap = 2, fcap = 2;
  while (cap < world * k) cap <<= 1;
  while (fcap < 2 * k) fcap <<= 1;
  const size_t smem = static_cast<size_t>(cap) * 16 + static_cast<size_t>((k + 1) & ~1) * 4 + static_cast<size_t>(k) * 24 + 16 +
                      static_cast<size_t>(fcap) * 8 + 64;
  B2R_CUDA(cudaFuncSetAttribute(hybrid_exchange_fuse_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024));
  hybrid_exchange_fuse_kernel<<<nq, 256, smem, static_cast<cudaStream_t>(stream)>>>(
      D, reinterpret_cast<const long long*>(I), S, reinterpret_cast<const long long*>(J), nq, k, x, F,
      reinterpret_cast<long long*>(O));
  B2R_CUDA(cudaGetLastError());
  return B2R_OK;
}
