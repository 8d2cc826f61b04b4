// hybrid_fuse.cuh — per-query device routines of the dense + BM25 fusion and of the multi-GPU exchange, shared by
// the stand-alone kernels (hybrid_fuse.cu) and by the BM25 kernel's fused tail (bm25.cu: the last CTA of a query
// fuses — or exchanges, merges and fuses — right after the scan has completed, without a kernel launch in between).
#pragma once

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

// Fused tail of the BM25 kernel (bm25.cu): what the last CTA of a query does with its BM25 top-k once the dense scan
// launched in front of it has completed.  mode 0: nothing (write S / J only); 1: fuse with the scan's D / I into F / O;
// 2: exchange + merge + fuse over the peers' buffers (x).  k <= kTailMaxK.
static const int kTailMaxK = 256;
struct HybridTail {
  int mode;
  int trigger_next;       // issue griddepcontrol.launch_dependents at the kernel's start: the NEXT step's scan
                          // (B2R_LAUNCH_OVERLAP_PREVIOUS) may queue up behind this step's scan right away
  const float* D;         // [nq, k] the scan's results (global ids)
  const long long* I;
  float* F;               // [nq, k]
  long long* O;
  ExchangeArgs x;
};

__host__ __device__ inline size_t exchange_smem_bytes(int world, int k) {
  int cap = 2, fcap = 2;
  while (cap < world * k) cap <<= 1;
  while (fcap < 2 * k) fcap <<= 1;
  return 16 + static_cast<size_t>(cap) * 16 + static_cast<size_t>((k + 1) & ~1) * 4 + static_cast<size_t>(k) * 24 + 16 +
         static_cast<size_t>(fcap) * 8 + 64;
}

// One query's exchange + merge + fusion (whole-CTA routine, 256 threads): Dq / Iq / Sq / Jq are this rank's k hits of
// query q (global or shared memory), Fq / Oq receive the k fused results.  smem_raw: scratch of exchange_smem_bytes().
__device__ __forceinline__ void exchange_merge_fuse_query(const float* Dq, const long long* Iq, const double* Sq,
                                                          const long long* Jq, int q, int k, const ExchangeArgs& x,
                                                          float* Fq, long long* Oq, unsigned char* smem_raw) {
  const int tid = threadIdx.x, nt = blockDim.x;
  int& timed_out = *reinterpret_cast<int*>(smem_raw);
  smem_raw += 16;
  if (tid == 0) timed_out = 0;
  __syncthreads();
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
      pD[i] = Dq[i];
      pI[i] = Iq[i];
      pS[i] = Sq[i];
      pJ[i] = Jq[i];
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
      Fq[i] = -3.4028235e38f;
      Oq[i] = -1;
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
  fuse_query(mD, mI, mS, mJ, k, k, k, Fq, Oq, nullptr, fbuf, fcap, red);
}

}  // namespace b2r

// BM25 top-k on `stream` (optionally beside the previous kernel) whose last CTA per query runs `tail` (csrc/bm25.cu).
// Returns B2R_ERR_UNSUPPORTED without launching anything when this (index, nq, k) cannot take the fused tail.
int bm25_topk_with_tail(b2r_bm25* h, const int32_t* term_ids_dev, const int32_t* term_offsets_dev, int nq, int k,
                        double* S_dev, int64_t* J_dev, int64_t id_offset, cudaStream_t stream, bool overlap_prev,
                        const b2r::HybridTail& tail);
