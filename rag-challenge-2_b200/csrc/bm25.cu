// bm25.cu — K4: BM25Okapi scoring over term-major CSR postings with a fused top-k.
//
// Reference semantics (rank-bm25 0.2.2, called at src/retrieval.py:261-265):
//   score[doc] = sum over query tokens t (duplicates kept, query order)
//                idf[t] * ( tf*(k1+1) / ( tf + k1*(1 - b + b*dl/avgdl) ) )           (float64)
//   top_n      = stable descending sort of ALL documents (zero-score docs are eligible; equal
//                scores keep ascending doc id).
//
// Layout in HBM: indptr int64[V+1]; doc_ids int32[P] ascending inside each term; weight[P]
// = tf*(k1+1)/(tf + k1*(1-b+b*dl/avgdl)) computed once at build time in float64 with exactly the
// operation order numpy uses, stored as float64 (12 B / posting, the default: scores equal rank_bm25's bit for bit) or,
// as an explicit opt-in, float32 (8 B / posting, within 1e-6 relative); idf float64[V].
//
// Kernel: doc-range partitioned.  CTA (x = doc block, y = query) owns `block_docs` consecutive
// documents with a float64 accumulator each in shared memory.  For every query token it binary-
// searches that token's posting list for its doc range (all searches of a chunk of tokens run in
// parallel), then streams the segment: acc[doc] += idf * w.  Tokens are applied one after the
// other (barrier in between), docs are unique within a token, so there are no atomics and the
// float64 sums are formed in the reference's order.  The select then scans the accumulators in
// shared memory: warps keep threshold lists of 128-bit keys (ordered float64 score, inverted id);
// a global per-query score threshold (atomicMax) lets later CTAs skip almost everything.  Per-CTA
// lists go to a global scratch and the last CTA of each query merges them.  The dense score vector
// is only written by the parity entry point b2r_bm25_scores_host.
//
// Algorithmic bytes per query: sum over its tokens of df_t * (4 + sizeof(weight)).
#include <algorithm>
#include <map>
#include <mutex>
#include <vector>

#include <string.h>

#include "b2r_common.cuh"
#include "hybrid_fuse.cuh"

namespace b2r {

static const int kBm25Warps = 8;
static const int kTermChunk = 64;

__device__ __forceinline__ u64 double_to_ordered(double s) {
  s = s + 0.0;
  const u64 u = static_cast<u64>(__double_as_longlong(s));
  return (u >> 63) ? ~u : (u | 0x8000000000000000ull);
}

__device__ __forceinline__ double ordered_to_double(u64 o) {
  const u64 u = (o >> 63) ? (o & 0x7FFFFFFFFFFFFFFFull) : ~o;
  return __longlong_as_double(static_cast<long long>(u));
}

__device__ __forceinline__ u128 shfl_u128(u128 v, int src) {
  u32 w0 = static_cast<u32>(v), w1 = static_cast<u32>(v >> 32), w2 = static_cast<u32>(v >> 64),
      w3 = static_cast<u32>(v >> 96);
  w0 = __shfl_sync(0xffffffffu, w0, src);
  w1 = __shfl_sync(0xffffffffu, w1, src);
  w2 = __shfl_sync(0xffffffffu, w2, src);
  w3 = __shfl_sync(0xffffffffu, w3, src);
  return (static_cast<u128>(w3) << 96) | (static_cast<u128>(w2) << 64) | (static_cast<u128>(w1) << 32) | w0;
}

// first position p in [lo, hi) with ids[p] >= target
__device__ __forceinline__ long long lower_bound_i32(const int* __restrict__ ids, long long lo, long long hi,
                                                     long long target) {
  while (lo < hi) {
    const long long mid = lo + ((hi - lo) >> 1);
    if (static_cast<long long>(__ldg(ids + mid)) < target) lo = mid + 1;
    else hi = mid;
  }
  return lo;
}

// A kernel launched BESIDE its predecessor (programmatic stream serialization) never needs that grid's data, but the
// kernels that follow in the stream (merge, fusion) rely on "after BM25" meaning "after the scan too".  Waiting on the
// previous grid just before exiting makes that ordering explicit (a no-op for an ordinary launch).
__device__ __forceinline__ void pdl_wait_for_previous_grid() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

template <class W>
struct Bm25Args {
  const long long* indptr;
  const int* doc_ids;
  const W* weight;
  const double* idf;
  long long n_docs;
  int block_docs;
  const int* term_ids;      // device, all queries back to back
  const int* term_offsets;  // device, [nq + 1]
  int q_base;               // first query of this launch (blockIdx.y is relative to it)
  int k, cap, n_sel_warps;
  u128* partial;  // [nq_launch][gridDim.x][k]
  u32* tickets;   // [nq_launch]
  u64* tau_g;     // [nq_launch] global score threshold (ordered float64 bits)
  double* S;      // [nq, k]
  long long* I;   // [nq, k]
  long long id_offset;
  double* dense;  // non-NULL: write scores of query q_base densely and skip the select
  int sel_cap;
  HybridTail tail;  // what the last CTA of a query does after the scan in front of this kernel has completed
};

template <class W>
__global__ void __launch_bounds__(kBm25Warps * 32) bm25_kernel(const Bm25Args<W> a) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  if (a.tail.trigger_next) asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  const int tid = threadIdx.x, nt = blockDim.x;
  const int warp = tid >> 5, lane = tid & 31;
  const int B = a.block_docs;
  double* acc = reinterpret_cast<double*>(smem_raw);                         // [B]
  long long* seg = reinterpret_cast<long long*>(smem_raw + sizeof(double) * B);  // [2 * kTermChunk]
  u64* ctl = reinterpret_cast<u64*>(seg + 2 * kTermChunk);                   // [4]
  u128* cand = reinterpret_cast<u128*>(ctl + 4);                             // [n_sel_warps][cap]

  const int ql = blockIdx.y;        // query index inside this launch
  const int qi = a.q_base + ql;     // query index in the batch
  const long long lo = static_cast<long long>(blockIdx.x) * B;
  const long long hi = min(a.n_docs, lo + B);
  const int nb = static_cast<int>(hi - lo);

  for (int i = tid; i < B; i += nt) acc[i] = 0.0;
  const int t0 = a.term_offsets[qi], t1 = a.term_offsets[qi + 1];
  for (int tc = t0; tc < t1; tc += kTermChunk) {
    const int nterm = min(kTermChunk, t1 - tc);
    __syncthreads();
    for (int j = tid; j < 2 * nterm; j += nt) {
      const int term = a.term_ids[tc + (j >> 1)];
      seg[j] = lower_bound_i32(a.doc_ids, a.indptr[term], a.indptr[term + 1], (j & 1) ? hi : lo);
    }
    __syncthreads();
    for (int j = 0; j < nterm; ++j) {
      const int term = a.term_ids[tc + j];
      const double idf = a.idf[term];
      const long long p0 = seg[2 * j], p1 = seg[2 * j + 1];
      // 4 independent (doc id, weight) loads in flight per thread before the dependent updates
      long long p = p0 + tid;
      for (; p + 3LL * nt < p1; p += 4LL * nt) {
        int dd[4];
        W ww[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          dd[u] = __ldg(a.doc_ids + p + static_cast<long long>(u) * nt);
          ww[u] = __ldg(a.weight + p + static_cast<long long>(u) * nt);
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const int dl = dd[u] - static_cast<int>(lo);
          // explicit rn intrinsics: no FMA contraction, numpy forms idf*w and the sum separately
          acc[dl] = __dadd_rn(acc[dl], __dmul_rn(idf, static_cast<double>(ww[u])));
        }
      }
      for (; p < p1; p += nt) {
        const int dl = __ldg(a.doc_ids + p) - static_cast<int>(lo);
        acc[dl] = __dadd_rn(acc[dl], __dmul_rn(idf, static_cast<double>(__ldg(a.weight + p))));
      }
      __syncthreads();
    }
  }
  __syncthreads();

  if (a.dense) {
    for (int i = tid; i < nb; i += nt) a.dense[lo + i] = acc[i];
    pdl_wait_for_previous_grid();
    return;
  }

  // ---------------- fused select over the accumulators ----------------
  const int cap = a.cap, k = a.k;
  if (warp < a.n_sel_warps) {
    u128* list = cand + static_cast<size_t>(warp) * cap;
    int cnt = 0;
    u128 tau = 0;
    volatile u64* tg = a.tau_g + ql;
    u64 taug = *tg;
    int since_refresh = 0;
    for (int base = warp * 32; base < nb; base += a.n_sel_warps * 32) {
      const int i = base + lane;
      const bool valid = i < nb;
      const u64 sk = valid ? double_to_ordered(acc[i]) : 0ull;
      const u128 key = valid ? ((static_cast<u128>(sk) << 64) | static_cast<u128>(0xFFFFFFFFu - static_cast<u32>(lo + i)))
                             : static_cast<u128>(0);
      unsigned m = __ballot_sync(0xffffffffu, valid && sk >= taug && key > tau);
      while (m) {
        const int src = __ffs(m) - 1;
        m &= m - 1;
        const u128 kk = shfl_u128(key, src);
        if (kk > tau) {
          if (lane == 0) list[cnt] = kk;
          ++cnt;
          if (cnt == cap) {
            __syncwarp();
            warp_sort_desc(list, cap, lane);
            cnt = k;
            tau = list[k - 1];
            __syncwarp();
            // k documents of this warp score >= tau: publish the bound, adopt the global one
            const u64 mine = static_cast<u64>(tau >> 64);
            if (lane == 0) atomicMax(const_cast<u64*>(a.tau_g + ql), mine);
            const u64 g = *tg;
            taug = g > mine ? g : mine;
          }
        }
      }
      if (++since_refresh == 8) {
        since_refresh = 0;
        const u64 g = *tg;
        taug = g > taug ? g : taug;
      }
    }
    for (int i = cnt + lane; i < cap; i += 32) list[i] = 0;
  }
  __syncthreads();

  // CTA merge -> per-CTA partial list
  const int region = a.n_sel_warps * cap;
  block_sort_desc(cand, region);
  u128* dst = a.partial + (static_cast<size_t>(ql) * gridDim.x + blockIdx.x) * k;
  for (int i = tid; i < k; i += nt) dst[i] = i < region ? cand[i] : static_cast<u128>(0);
  __threadfence();
  __syncthreads();
  u32* s_last = reinterpret_cast<u32*>(ctl + 2);
  if (tid == 0) {
    const u32 t = atomicAdd(a.tickets + ql, 1u);
    *s_last = (t == gridDim.x - 1) ? 1u : 0u;
  }
  __syncthreads();
  if (*s_last == 0) return;
  __threadfence();
  // last CTA of this query: the accumulator area becomes the select scratch
  u128* sel = reinterpret_cast<u128*>(smem_raw);
  block_select_lists(a.partial + static_cast<size_t>(ql) * gridDim.x * k, gridDim.x, k, k, sel, a.sel_cap, ctl);
  if (a.tail.mode != 0) {
    // ---- fused tail (k <= 256): BM25 top-k -> shared memory, wait for the scan, then fuse / exchange in this CTA ----
    const u128 key = tid < k ? sel[tid] : static_cast<u128>(0);
    __syncthreads();
    double* mS = reinterpret_cast<double*>(smem_raw);               // [256]
    long long* mJ = reinterpret_cast<long long*>(smem_raw + 2048);  // [256]
    float* mD = reinterpret_cast<float*>(smem_raw + 4096);          // [256]
    long long* mI = reinterpret_cast<long long*>(smem_raw + 5120);  // [256]
    unsigned char* scratch = smem_raw + 7168;
    if (tid < k) {
      mS[tid] = key == 0 ? -1.7976931348623157e308 : ordered_to_double(static_cast<u64>(key >> 64));
      mJ[tid] = key == 0 ? -1 : static_cast<long long>(0xFFFFFFFFu - static_cast<u32>(key)) + a.id_offset;
    }
    if (tid == 0) {  // the control words are free again before anything of a chained next step can touch them
      a.tickets[ql] = 0;
      a.tau_g[ql] = 0;
    }
    // The merge above ran while the scan was still streaming; from here on the scan's results are needed.
    pdl_wait_for_previous_grid();
    const size_t o = static_cast<size_t>(qi) * k;
    if (tid < k) {
      if (a.S) {
        a.S[o + tid] = mS[tid];
        a.I[o + tid] = mJ[tid];
      }
      mD[tid] = __ldcg(a.tail.D + o + tid);
      mI[tid] = __ldcg(a.tail.I + o + tid);
    }
    __syncthreads();
    if (a.tail.mode == 1) {
      int fcap = 2;
      while (fcap < 2 * k) fcap <<= 1;
      fuse_query(mD, mI, mS, mJ, k, k, k, a.tail.F + o, a.tail.O + o, nullptr, reinterpret_cast<u64*>(scratch), fcap,
                 reinterpret_cast<double*>(scratch + static_cast<size_t>(fcap) * 8));
    } else {
      exchange_merge_fuse_query(mD, mI, mS, mJ, qi, k, a.tail.x, a.tail.F + o, a.tail.O + o, scratch);
    }
    return;
  }
  // Launched beside the scan: the merge above ran while the scan was still streaming; only the publication of the
  // result waits for it, so that "after BM25" means "after the scan too" for the kernels that follow, and so that
  // S / I may alias buffers the previous step's fusion kernel is still reading when steps are chained.
  pdl_wait_for_previous_grid();
  for (int i = tid; i < k; i += nt) {
    const u128 key = sel[i];
    const size_t o = static_cast<size_t>(qi) * k + i;
    if (key == 0) {
      a.S[o] = -1.7976931348623157e308;
      a.I[o] = -1;
    } else {
      a.S[o] = ordered_to_double(static_cast<u64>(key >> 64));
      a.I[o] = static_cast<long long>(0xFFFFFFFFu - static_cast<u32>(key)) + a.id_offset;
    }
  }
  if (tid == 0) {
    a.tickets[ql] = 0;
    a.tau_g[ql] = 0;
  }
}


// mbarrier wait with a watchdog: a pipeline bug traps instead of hanging the GPU
__device__ __forceinline__ void mbar_wait_wd(u64* bar, u32 parity) {
  const u32 addr = smem_u32(bar);
  for (u32 spins = 0;; ++spins) {
    u32 done;
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t"
        "}"
        : "=r"(done)
        : "r"(addr), "r"(parity)
        : "memory");
    if (done) return;
    if (spins == (1u << 27)) __trap();
  }
}

// one NON-BLOCKING probe of an mbarrier phase (test_wait; try_wait may suspend the thread for a while): true when the
// phase with this parity has completed
__device__ __forceinline__ bool mbar_probe(u64* bar, u32 parity) {
  u32 done;
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t"
      "}"
      : "=r"(done)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return done != 0u;
}

// barrier over the 256 consumer threads of the streaming kernel (the producer warp never joins)
__device__ __forceinline__ void named_bar_sync_256() { asm volatile("bar.sync 1, 256;" ::: "memory"); }

// OR-reduction over the 256 consumer threads in ONE named barrier
__device__ __forceinline__ bool named_bar_red_or_256(bool pred) {
  u32 out;
  asm volatile(
      "{\n\t"
      ".reg .pred p, q;\n\t"
      "setp.ne.u32 q, %1, 0;\n\t"
      "barrier.red.or.pred p, 1, 256, q;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t"
      "}"
      : "=r"(out)
      : "r"(static_cast<u32>(pred))
      : "memory");
  return out != 0;
}

// bitonic sort (descending) by the 256 consumer threads; n is a power of two
template <class K>
__device__ __forceinline__ void consumer_sort_desc(K* a, int n, int ct) {
  for (int ls = 1; (1 << ls) <= n; ++ls) {
    for (int lj = ls - 1; lj >= 0; --lj) {
      const int stride = 1 << lj;
      for (int t = ct; t < (n >> 1); t += 256) {
        const int i = ((t >> lj) << (lj + 1)) | (t & (stride - 1));
        const bool up = ((i >> ls) & 1) != 0;
        cmpxchg_desc(a, i, i | stride, up);
      }
      named_bar_sync_256();
    }
  }
}

// ----------------------------------------------------------------------------------------------
// K4s: streaming BM25 kernel (default).  Same arithmetic and results as bm25_kernel above; organised like
// the flat scan: persistent CTAs, one producer warp driving the bulk-copy engine, eight consumer warps.
//
//   work item = (query, block of kStreamBlk = 4096 documents); CTA c takes items c, c + grid, ...
//   producer   per item and token: segment bounds from the block index skip[term][block] (built once per
//              index; no binary search on the query path), then 16-byte aligned bulk copies of the segment's
//              (doc id, weight) arrays into a ring of 1024-posting stages — the stages of a token group are
//              numbered by a warp scan and issued a ring round at a time, one per lane; a token without postings
//              in the block issues nothing.  It runs ahead of the consumers across item boundaries, so the
//              index / term-table latency is hidden behind their work.
//   consumers  acc[doc] += idf * w for the postings of a stage (documents are unique within a token, tokens are
//              separated by a named barrier: float64 sums in the reference's order, no atomics), then the
//              fused select over the 4096 accumulators: candidates are the documents with score >= the query's
//              global threshold tau_g, reduced to the best k by a histogram select on the ordered score bits
//              (equal scores: ascending id), sorted, written to the per-(query, block) list; bm25_merge_kernel
//              (one CTA per query) merges the lists.  tau_g = the minimum over the query's k threshold slots
//              (slot j: best score seen among documents with id % k == j — k distinct documents, hence a lower
//              bound of the k-th best score); batches run a pilot launch over every query's first blocks so that
//              the main launch starts with a threshold (bm25_run_stream).
// Shared memory: 32 KiB accumulators (the select's 8 KiB histogram + 8 KiB list borrow their first 16 KiB) + the ring;
// float32 weights: 5 stages x 1024 postings x 8 B = 40 KiB -> 73 KiB; float64 weights: 3 x 1024 x 12 B = 36 KiB -> 69 KiB;
// three CTAs per SM either way (see bm25_stream_launch).
// Algorithmic bytes per query: sum_t df_t * (4 + sizeof(weight)); the block index adds 8 B per (token, block).
// ----------------------------------------------------------------------------------------------
static const int kStreamBlk = 4096;       // documents per work item (block index granularity)
// ring geometry (postings per stage, stages): template parameters of the kernel, chosen in bm25_stream_launch
static const int kStreamSelCap = 512;     // candidate list (u128 keys)
static const int kStreamMaxK = 384;       // k + 64 spill keys must fit the list
static const int kStreamThreads = 9 * 32; // 8 consumer warps + producer warp

template <class W>
struct Bm25StreamArgs {
  const long long* indptr;
  const int* doc_ids;
  const W* weight;
  const double* idf;
  const u32* skip;          // [V][n_blocks + 1] first posting (relative to indptr[t]) with doc >= b * kStreamBlk; may be NULL
  long long n_docs;
  int n_blocks;
  int b_begin, b_count;     // this launch covers blocks [b_begin, b_begin + b_count) of every query (pilot / main launch)
  const int* term_ids;
  const int* term_offsets;
  int q_base, nq_launch;
  int k;
  u128* partial;            // [nq_launch][n_blocks][k]
  u64* tau_g;               // [nq_launch]
  u64* slots;               // [nq_launch][k] best score seen among the documents with id % k == slot (0: none); may be NULL
  double* dense;            // non-NULL: write the scores of query q_base densely, skip the select
  int debug_skip;           // perf experiments only (wrong results): 1 no select, 2 no accumulate
};

struct StreamMeta {
  double idf;
  int fv, lv;                // valid elements of the stage: [fv, lv) (the copy window is 16-byte aligned, hence wider)
  int flags;                 // 1: last stage of the token, 2: last stage of the item
  int pad;
};

template <class W, int CHUNK, int STAGES, int MINB>
__global__ void __launch_bounds__(kStreamThreads, MINB) bm25_stream_kernel(const Bm25StreamArgs<W> a) {
  extern __shared__ __align__(128) unsigned char stream_smem[];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  constexpr int kStageBytes = CHUNK * (4 + static_cast<int>(sizeof(W)));
  double* acc = reinterpret_cast<double*>(stream_smem);                                   // [kStreamBlk]
  unsigned char* ring = stream_smem + kStreamBlk * sizeof(double);                        // [stages][ids | weights]
  StreamMeta* meta = reinterpret_cast<StreamMeta*>(ring + STAGES * kStageBytes);  // [stages]
  u64* full = reinterpret_cast<u64*>(meta + STAGES);
  u64* empty = full + STAGES;
  u64* ctl = empty + STAGES;                                                    // [8] select control words
  // The select's histogram and candidate list live IN the accumulator area (its first 16 KiB): when a block has a
  // candidate at all, every thread has moved its accumulators to registers and cleared them, and nothing touches the
  // area until the next item's first stage.  A block that used them hands the area back zeroed.  The 16 KiB saved are
  // ring: 1024-posting stages (fewer hand-overs per item, which is what the kernel's time follows) at three CTAs per SM.
  static_assert(kSelBins * 4 + kStreamSelCap * 16 <= kStreamBlk * 8, "select scratch must fit the accumulator area");
  u32* hist = reinterpret_cast<u32*>(acc);                                             // [kSelBins]
  u128* sel = reinterpret_cast<u128*>(stream_smem + kSelBins * 4);                     // [kStreamSelCap]
  constexpr int kScratchRows = (kSelBins * 4 + kStreamSelCap * 16) / 8 / 256;          // accumulator rows under them

  if (tid == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], 8);
    }
    mbar_fence_init();
  }
  __syncthreads();
  const long long n_items = static_cast<long long>(a.nq_launch) * a.b_count;
  // CTA c takes items c, c + grid, ...: at any time the whole grid works on one or two queries, so a query's
  // threshold tau_g saturates within its first wave of blocks and the remaining blocks select next to nothing
  const long long item_lo = blockIdx.x, item_hi = n_items, item_step = gridDim.x;

  if (warp == 8) {
    // ======================= producer warp =======================
    // The token table of the current query (term, list start, idf) stays in registers from item to item; only
    // the two block-index entries per token change, and those of the NEXT item are loaded while this item's
    // copies are being issued.
    const u64 policy = l2_policy_evict_first();
    u32 gs = 0;  // stages issued so far by this CTA: stage n lives in slot n % STAGES, round parity (n / STAGES) & 1
    int cur_q = -1, t0 = 0, t1 = 0;
    int term = 0;          // lane-held token table (first 32 tokens of the query)
    long long ib = 0, ie = 0;
    double idfv = 0.0;
    u32 nx0 = 0, nx1 = 0;  // prefetched block-index entries of the next item (lane-held)
    bool have_next = false;
    for (long long item = item_lo; item < item_hi; item += item_step) {
      const int ql = static_cast<int>(item / a.b_count);
      const int b = a.b_begin + static_cast<int>(item - static_cast<long long>(ql) * a.b_count);
      const long long lo = static_cast<long long>(b) * kStreamBlk;
      const long long hi = min(a.n_docs, lo + kStreamBlk);
      if (ql != cur_q) {
        cur_q = ql;
        const int qi = a.q_base + ql;
        t0 = a.term_offsets[qi];
        t1 = a.term_offsets[qi + 1];
        if (t0 + lane < t1) {
          term = a.term_ids[t0 + lane];
          ib = a.indptr[term];
          ie = a.indptr[term + 1];
          idfv = a.idf[term];
        }
        have_next = false;
      }
      if (t1 == t0) {  // a query without known tokens: one empty stage closes the item
        if (lane == 0) {
          const int slot = static_cast<int>(gs % STAGES);
          mbar_wait_wd(&empty[slot], ((gs / STAGES) & 1u) ^ 1u);
          meta[slot].fv = meta[slot].lv = 0;
          meta[slot].idf = 0.0;
          meta[slot].flags = 3;
          mbar_arrive(&full[slot]);
        }
        ++gs;
        continue;
      }
      for (int tb = t0; tb < t1; tb += 32) {
        const int j = tb + lane;
        long long p0 = 0, p1 = 0;
        double tidf = 0.0;
        if (j < t1) {
          int tt = term;
          long long tib = ib, tie = ie;
          tidf = idfv;
          if (tb != t0) {  // queries longer than 32 tokens: later rounds are fetched on the fly
            tt = a.term_ids[j];
            tib = a.indptr[tt];
            tie = a.indptr[tt + 1];
            tidf = a.idf[tt];
          }
          if (a.skip) {
            u32 e0, e1;
            if (tb == t0 && have_next) {
              e0 = nx0;
              e1 = nx1;
            } else {
              const u32* row = a.skip + static_cast<size_t>(tt) * (a.n_blocks + 1) + b;
              e0 = row[0];
              e1 = row[1];
            }
            p0 = tib + e0;
            p1 = tib + e1;
          } else {
            p0 = lower_bound_i32(a.doc_ids, tib, tie, lo);
            p1 = lower_bound_i32(a.doc_ids, p0, tie, hi);
          }
        }
        if (tb == t0) {  // prefetch the next item's entries when it belongs to the same query
          have_next = false;
          if (a.skip && item + item_step < item_hi && b + item_step < a.b_begin + a.b_count) {
            if (t0 + lane < t1) {
              const u32* row = a.skip + static_cast<size_t>(term) * (a.n_blocks + 1) + (b + item_step);
              nx0 = row[0];
              nx1 = row[1];
            }
            have_next = true;
          }
        }
        const int nt = min(32, t1 - tb);
        // One lane issuing every stage (wait for the slot, four stores, expect-tx, two bulk copies: several hundred
        // cycles of dependent instructions) was the critical path of the whole kernel — its time scaled with the
        // number of producers per SM, not with the postings.  So the stages of a token group are flattened in token
        // order (lane u owns token u's aligned window [a0, a1) and its chunk count; an exclusive scan numbers them)
        // and issued STAGES at a time, one per lane, in parallel: the slots of a round are distinct, and each lane
        // waits only for its own slot's previous use, which every earlier round has already issued.
        const long long wa0 = p0 & ~3LL, wa1 = (p1 + 3) & ~3LL;  // 16-byte aligned copy window; consumers mask by index
        // a token without postings in this block issues nothing (rare terms miss most blocks; every stage is a visit
        // and a token barrier for the eight consumer warps)
        const int nch = (lane < nt && p1 > p0) ? static_cast<int>((wa1 - wa0 + CHUNK - 1) / CHUNK) : 0;
        int incl = nch;
#pragma unroll
        for (int mm = 1; mm < 32; mm <<= 1) {
          const int v = __shfl_up_sync(0xffffffffu, incl, mm);
          if (lane >= mm) incl += v;
        }
        const int pre = incl - nch;
        const int n_stages = __shfl_sync(0xffffffffu, incl, 31);
        const bool last_group = tb + 32 >= t1;
        if (n_stages == 0) {
          if (last_group) {  // nothing of the last tokens in this block: a bare marker closes the item
            if (lane == 0) {
              const int slot = static_cast<int>(gs % STAGES);
              mbar_wait_wd(&empty[slot], ((gs / STAGES) & 1u) ^ 1u);
              meta[slot].fv = meta[slot].lv = 0;
              meta[slot].idf = 0.0;
              meta[slot].flags = 3;
              mbar_arrive(&full[slot]);
            }
            ++gs;
          }
          continue;
        }
        for (int base = 0; base < n_stages; base += STAGES) {
          const int g = base + lane;
          const bool on = lane < STAGES && g < n_stages;
          int u = 0;  // the token of stage g: the largest u < nt with pre_u <= g (tokens after the owner start above g)
#pragma unroll
          for (int bit = 16; bit > 0; bit >>= 1) {
            const int cand = u + bit;
            const int pc = __shfl_sync(0xffffffffu, pre, cand & 31);
            if (cand < nt && pc <= g) u = cand;
          }
          const long long q0 = __shfl_sync(0xffffffffu, p0, u), q1 = __shfl_sync(0xffffffffu, p1, u);
          const double qidf = __shfl_sync(0xffffffffu, tidf, u);
          const int pu = __shfl_sync(0xffffffffu, pre, u), nu = __shfl_sync(0xffffffffu, nch, u);
          const long long a0 = q0 & ~3LL, a1 = (q1 + 3) & ~3LL;
          const int c = g - pu;
          const long long c0 = a0 + static_cast<long long>(c) * CHUNK;
          const long long c1 = min(a1, c0 + CHUNK);
          const int cnt = static_cast<int>(c1 - c0);
          const bool last_chunk = c == nu - 1;
          const bool last_token = last_group && g == n_stages - 1;  // the item's last stage (later tokens are absent here)
          const u32 n = gs + static_cast<u32>(g);
          const int slot = static_cast<int>(n % STAGES);
          const u32 free_parity = ((n / STAGES) & 1u) ^ 1u;
          // A lane must not WAIT for its slot inside the branch: the warp would reconverge behind the slowest lane, i.e.
          // issue nothing of a round until every slot of the ring is free — the ring would drain between rounds.  Every
          // lane probes its slot once per turn instead; slots come free in order, so the round's stages go out one by
          // one as the consumers release them, and the lanes still issue in parallel whenever several are free.
          bool todo = on;
          for (u32 spins = 0; __any_sync(0xffffffffu, todo); ++spins) {
            if (spins == (1u << 27)) __trap();
            if (!(todo && mbar_probe(&empty[slot], free_parity))) continue;
            todo = false;
            StreamMeta& m = meta[slot];
            m.idf = qidf;
            m.fv = static_cast<int>(max(q0 - c0, 0LL));
            m.lv = static_cast<int>(min(q1 - c0, static_cast<long long>(cnt)));
            m.flags = (last_chunk ? 1 : 0) | ((last_chunk && last_token) ? 2 : 0);
            unsigned char* st = ring + static_cast<size_t>(slot) * kStageBytes;
            if (cnt > 0) {
              mbar_arrive_expect_tx(&full[slot], static_cast<u32>(cnt) * (4u + static_cast<u32>(sizeof(W))));
              bulk_g2s(st, a.doc_ids + c0, static_cast<u32>(cnt) * 4u, &full[slot], policy);
              bulk_g2s(st + CHUNK * 4, a.weight + c0, static_cast<u32>(cnt) * static_cast<u32>(sizeof(W)), &full[slot],
                       policy);
            } else {
              mbar_arrive(&full[slot]);
            }
          }
          __syncwarp();
        }
        gs += static_cast<u32>(n_stages);
      }
    }
    pdl_wait_for_previous_grid();
    return;
  }

  // ======================= consumer warps =======================
  const int ct = tid;  // 0..255
#pragma unroll
  for (int i = 0; i < kStreamBlk / 256; ++i) acc[i * 256 + ct] = 0.0;
  named_bar_sync_256();
  int s = 0;
  u32 ph = 0;
  u32* s_cnt = reinterpret_cast<u32*>(ctl);        // [0] candidates, [1] selected, [2] threshold bin, [3] above
  u64* s_mn = ctl + 2;
  u64* s_mx = ctl + 3;
  // item -> (query, block) without a 64-bit division per item: the coordinates advance with the stride
  int ql = static_cast<int>(item_lo / a.b_count);
  int b = a.b_begin + static_cast<int>(item_lo - static_cast<long long>(ql) * a.b_count);
  const int step_q = static_cast<int>(item_step / a.b_count);
  const int step_b = static_cast<int>(item_step - static_cast<long long>(step_q) * a.b_count);
  const bool no_acc = (a.debug_skip & 2) != 0;
  for (long long item = item_lo; item < item_hi; item += item_step) {
    const int lo_i = b * kStreamBlk;
    const long long lo = lo_i;
    const int nb = static_cast<int>(min(a.n_docs - lo, static_cast<long long>(kStreamBlk)));
    // the query's threshold is fetched now and used after the accumulation (its latency hides behind the stages; a
    // slightly stale value is only a lower bound that is a little lower)
    const u64 taug_early = a.dense ? 0ull : *(reinterpret_cast<volatile u64*>(a.tau_g) + ql);
    // ---- accumulate: stages arrive in token order (acc is zero: cleared at start / by the previous select) ----
    for (;;) {
      mbar_wait_wd(&full[s], ph);
      const double idf = meta[s].idf;
      const int fv = meta[s].fv, lv = no_acc ? 0 : meta[s].lv, fl = meta[s].flags;
      const int* ids = reinterpret_cast<const int*>(ring + static_cast<size_t>(s) * kStageBytes);
      const W* ws = reinterpret_cast<const W*>(ring + static_cast<size_t>(s) * kStageBytes + CHUNK * 4);
      for (int i = fv + ct; i < lv; i += 256) {
        const int dl = ids[i] - lo_i;
        // explicit rn intrinsics: no FMA contraction, numpy forms idf*w and the sum separately
        acc[dl] = __dadd_rn(acc[dl], __dmul_rn(idf, static_cast<double>(ws[i])));
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&empty[s]);
      if (++s == STAGES) {
        s = 0;
        ph ^= 1u;
      }
      if (fl & 1) named_bar_sync_256();  // token boundary: the next token may touch the same documents
      if (fl & 2) break;
    }
    const int ql_now = ql, b_now = b;  // advance the coordinates for the next item (every path below may `continue`)
    b += step_b;
    ql += step_q;
    if (b >= a.b_begin + a.b_count) {
      b -= a.b_count;
      ++ql;
    }
    if (a.dense) {
      for (int i = ct; i < kStreamBlk; i += 256) {
        if (i < nb) a.dense[lo + i] = acc[i];
        acc[i] = 0.0;
      }
      named_bar_sync_256();
      continue;
    }

    // ---- fused select: best k of this block among the documents that can still matter ----
    // Fast path (the usual case after a query's first wave): compare the float64 scores with the threshold AS DOUBLES —
    // three instructions per document; the ordered-bit keys are only formed when some document of the block survives.
    const int k = a.k;
    const u64 taug = (a.debug_skip & 1) ? ~0ull : taug_early;
    // ordered bits -> double is monotonic; 0 means "no threshold yet" (below every score), ~0 is above every score
    const double tau_d = taug == 0ull ? -__longlong_as_double(0x7FF0000000000000LL)
                                      : (taug == ~0ull ? __longlong_as_double(0x7FF0000000000000LL) : ordered_to_double(taug));
    double sv[kStreamBlk / 256];
    u32 act = 0;  // bit i: document i * 256 + ct is a live candidate
#pragma unroll
    for (int i = 0; i < kStreamBlk / 256; ++i) {
      const int dcl = i * 256 + ct;
      sv[i] = acc[dcl];
      acc[dcl] = 0.0;  // ready for the next item (this thread is the only reader of these entries)
      if (sv[i] >= tau_d) act |= 1u << i;
    }
    if (nb < kStreamBlk) {  // the corpus ends inside this block
#pragma unroll
      for (int i = 0; i < kStreamBlk / 256; ++i) {
        if (i * 256 + ct >= nb) act &= ~(1u << i);
      }
    }
    if (taug == ~0ull) act = 0;
    u128* dst = a.partial + (static_cast<size_t>(ql_now) * a.n_blocks + b_now) * k;
    // ONE barrier decides the usual case (after the first wave of a query no document of a block reaches tau_g): it
    // also orders this item's accumulator reads / clears before the next item's updates
    if (!named_bar_red_or_256(act != 0u)) {
      for (int i = ct; i < k; i += 256) dst[i] = 0;
      continue;
    }
    u64 sk[kStreamBlk / 256];
#pragma unroll
    for (int i = 0; i < kStreamBlk / 256; ++i) sk[i] = double_to_ordered(sv[i]);
    if (ct == 0) {
      s_cnt[0] = 0;
      s_cnt[1] = 0;
    }
    named_bar_sync_256();
    {
      const u32 c = __reduce_add_sync(0xffffffffu, static_cast<u32>(__popc(act)));
      if (lane == 0 && c) atomicAdd(&s_cnt[0], c);
    }
    named_bar_sync_256();
    int need = min(k, static_cast<int>(s_cnt[0]));
    int live = static_cast<int>(s_cnt[0]);
    // narrow the live set by histogram levels until what is left fits the list
    while (live > need && live > 64) {
      if (ct == 0) {
        *s_mn = ~0ull;
        *s_mx = 0ull;
      }
      for (int i = ct; i < kSelBins; i += 256) hist[i] = 0;
      named_bar_sync_256();
      u64 mn = ~0ull, mx = 0ull;
#pragma unroll
      for (int i = 0; i < kStreamBlk / 256; ++i) {
        if ((act >> i) & 1u) {
          mn = sk[i] < mn ? sk[i] : mn;
          mx = sk[i] > mx ? sk[i] : mx;
        }
      }
      // warp-reduce first: 64-bit shared-memory min / max are CAS loops and collapse under 256-way contention
#pragma unroll
      for (int mm = 16; mm > 0; mm >>= 1) {
        const u64 omn = __shfl_xor_sync(0xffffffffu, mn, mm), omx = __shfl_xor_sync(0xffffffffu, mx, mm);
        mn = omn < mn ? omn : mn;
        mx = omx > mx ? omx : mx;
      }
      if (lane == 0 && mn <= mx) {
        atomicMin(s_mn, mn);
        atomicMax(s_mx, mx);
      }
      named_bar_sync_256();
      const u64 gmn = *s_mn, range = *s_mx - gmn;
      if (range == 0) break;  // all live scores equal: ascending id decides (below)
      int sh = 0;
      while ((range >> sh) >= static_cast<u64>(kSelBins)) ++sh;
      // scores cluster in a few bins: lanes with the same bin combine their increments (one atomic per group)
#pragma unroll
      for (int i = 0; i < kStreamBlk / 256; ++i) {
        const bool on = (act >> i) & 1u;
        const u32 bin = on ? static_cast<u32>((sk[i] - gmn) >> sh) : 0xFFFFFFFFu;
        const u32 peers = __match_any_sync(0xffffffffu, bin);
        if (on && lane == __ffs(peers) - 1) atomicAdd(&hist[bin], static_cast<u32>(__popc(peers)));
      }
      named_bar_sync_256();
      if (warp == 0) {  // bins from the top until `need` keys are covered
        constexpr int per = kSelBins / 32;
        const int top = kSelBins - per * lane;
        u32 sum = 0;
        for (int bb = top - 1; bb >= top - per; --bb) sum += hist[bb];
        u32 incl = sum;
#pragma unroll
        for (int mm = 1; mm < 32; mm <<= 1) {
          const u32 v = __shfl_up_sync(0xffffffffu, incl, mm);
          if (lane >= mm) incl += v;
        }
        const u32 before = incl - sum;
        if (before < static_cast<u32>(need) && incl >= static_cast<u32>(need)) {
          u32 accn = before;
          int bb = top - 1;
          for (; bb > top - per; --bb) {
            if (accn + hist[bb] >= static_cast<u32>(need)) break;
            accn += hist[bb];
          }
          s_cnt[2] = static_cast<u32>(bb);
          s_cnt[3] = accn;  // keys in the bins above the threshold bin
        }
      }
      named_bar_sync_256();
      const u32 bstar = s_cnt[2], above = s_cnt[3], bucket = hist[bstar];
      // keys above the threshold bin are in the answer; the bin itself stays live
#pragma unroll
      for (int i = 0; i < kStreamBlk / 256; ++i) {
        if ((act >> i) & 1u) {
          const u32 bin = static_cast<u32>((sk[i] - gmn) >> sh);
          if (bin > bstar) {
            const u32 pos = atomicAdd(&s_cnt[1], 1u);
            sel[pos] = (static_cast<u128>(sk[i]) << 64) | static_cast<u128>(0xFFFFFFFFu - static_cast<u32>(lo + i * 256 + ct));
          }
          if (bin != bstar) act &= ~(1u << i);
        }
      }
      need -= static_cast<int>(above);
      live = static_cast<int>(bucket);
      named_bar_sync_256();
    }
    if (live > need && live > 64) {
      // all live documents share one score: the `need` lowest ids win.  Rank = live documents before this one
      // in id order (row i * 256 + ct): rows in order, threads in order inside a row.
#pragma unroll
      for (int i = 0; i < kStreamBlk / 256; ++i) {
        if (need <= 0) break;  // uniform
        const bool on = (act >> i) & 1u;
        const u32 bal = __ballot_sync(0xffffffffu, on);
        u32* wcnt = hist;  // [8] per-warp counts of this row
        if (lane == 0) wcnt[warp] = __popc(bal);
        named_bar_sync_256();
        u32 before = 0, row_total = 0;
        for (int w = 0; w < 8; ++w) {
          const u32 c = wcnt[w];
          if (w < warp) before += c;
          row_total += c;
        }
        const u32 rank = before + __popc(bal & ((1u << lane) - 1u));
        if (on && rank < static_cast<u32>(need)) {
          const u32 pos = atomicAdd(&s_cnt[1], 1u);
          sel[pos] = (static_cast<u128>(sk[i]) << 64) | static_cast<u128>(0xFFFFFFFFu - static_cast<u32>(lo + i * 256 + ct));
        }
        need -= static_cast<int>(min(row_total, static_cast<u32>(need)));
        named_bar_sync_256();
      }
      act = 0;
    }
    // what is still live (<= need, or <= 64 spill keys) joins the list; the sort settles the order
#pragma unroll
    for (int i = 0; i < kStreamBlk / 256; ++i) {
      if ((act >> i) & 1u) {
        const u32 pos = atomicAdd(&s_cnt[1], 1u);
        sel[pos] = (static_cast<u128>(sk[i]) << 64) | static_cast<u128>(0xFFFFFFFFu - static_cast<u32>(lo + i * 256 + ct));
      }
    }
    named_bar_sync_256();
    const int n_sel = static_cast<int>(s_cnt[1]);
    int p2 = 2;
    while (p2 < n_sel) p2 <<= 1;
    if (p2 > kStreamSelCap) p2 = kStreamSelCap;
    for (int i = n_sel + ct; i < p2; i += 256) sel[i] = 0;
    named_bar_sync_256();
    if (p2 <= 128) {  // short list: one warp sorts it without block-wide barriers
      if (warp == 0) warp_sort_desc(sel, p2, lane);
      named_bar_sync_256();
    } else {
      consumer_sort_desc(sel, p2, ct);
    }
    for (int i = ct; i < k; i += 256) dst[i] = i < p2 ? sel[i] : static_cast<u128>(0);
    if (ct == 0 && n_sel >= k) {
      // k documents of this block score >= sel[k - 1]: nothing below it can enter the query's top k
      atomicMax(a.tau_g + ql_now, static_cast<u64>(sel[k - 1] >> 64));
    }
    // A block's own k-th best is a weak bound (the 30th of 4096 documents is far below the 30th of ten million).  The
    // query also keeps k slots, slot j = the best score seen among the documents with id % k == j: the slots hold k
    // DISTINCT documents, so their minimum is a lower bound of the query's k-th best score — and since the best
    // documents spread over the slots it settles near the (k ln k)-th best score seen so far.  With it, most blocks
    // after a query's first wave have no document at or above the threshold and leave through the one-barrier path.
    if (a.slots != nullptr) {
      u64* sl = a.slots + static_cast<size_t>(ql_now) * k;
      const int m = min(min(n_sel, p2), k);
      for (int i = ct; i < m; i += 256) {
        const u128 key = sel[i];
        if (key != 0) {
          u64* slot = sl + (0xFFFFFFFFu - static_cast<u32>(key)) % static_cast<u32>(k);
          const u64 sc = static_cast<u64>(key >> 64);
          if (sc > *reinterpret_cast<volatile u64*>(slot)) atomicMax(slot, sc);
        }
      }
    }
    named_bar_sync_256();  // every read of the list is done (and the slot updates of this block)
#pragma unroll
    for (int i = 0; i < kScratchRows; ++i) acc[i * 256 + ct] = 0.0;  // the scratch was accumulator space: zero it again
    named_bar_sync_256();  // ... before any warp starts the next item
    if (a.slots != nullptr && warp == 0) {
      const u64* sl = a.slots + static_cast<size_t>(ql_now) * k;
      u64 mn = ~0ull;
      for (int j = lane; j < k; j += 32) {
        const u64 v = *reinterpret_cast<const volatile u64*>(sl + j);
        mn = v < mn ? v : mn;
      }
#pragma unroll
      for (int mm = 16; mm > 0; mm >>= 1) {
        const u64 o = __shfl_xor_sync(0xffffffffu, mn, mm);
        mn = o < mn ? o : mn;
      }
      if (lane == 0 && mn != 0ull) atomicMax(a.tau_g + ql_now, mn);  // 0: some slot is still empty
    }
  }
  pdl_wait_for_previous_grid();
}

// one CTA per query: merge the per-block lists of the streaming kernel into the query's top k, reset its threshold.
// theta = k-th largest of the lists' best scores (k lists hold a key >= theta each, so the answer lies above it);
// the keys above theta — a little over k of them — are gathered and the best k selected, both by histogram.
static const int kMergeLists = 8192;   // list maxima the threshold step can take (33M documents per shard)
static const int kMergeCap = 4096;     // gathered keys

__global__ void __launch_bounds__(1024) bm25_merge_kernel(const u128* __restrict__ partial, int n_lists, int k, int q_base,
                                                          u64* __restrict__ tau_g, double* __restrict__ S,
                                                          long long* __restrict__ I, long long id_offset) {
  extern __shared__ __align__(16) unsigned char merge_smem[];
  u128* buf = reinterpret_cast<u128*>(merge_smem);                       // [kMergeCap]; first the u32 list maxima
  u128* top = buf + kMergeCap;                                           // [1024]
  u32* hist = reinterpret_cast<u32*>(top + 1024);                        // [kSelBins + 4]
  __shared__ u64 ctl[4];
  __shared__ u32 s_n;
  const int ql = blockIdx.x, qi = q_base + ql, tid = threadIdx.x, nt = blockDim.x;
  const u128* src = partial + static_cast<size_t>(ql) * n_lists * k;
  const long long total = static_cast<long long>(n_lists) * k;
  u32 theta = 0;
  if (n_lists >= k && n_lists <= kMergeLists) {
    u32* mx = reinterpret_cast<u32*>(buf);
    u32* top32 = reinterpret_cast<u32*>(top);
    for (int l = tid; l < n_lists; l += nt) mx[l] = key_score_bits(src[static_cast<long long>(l) * k]);
    __syncthreads();
    block_topk_hist(mx, n_lists, kMergeLists, k, hist, top32, 1024);
    theta = top32[k - 1];
    __syncthreads();
  }
  if (tid == 0) s_n = 0;
  __syncthreads();
  for (long long i = tid; i < total; i += nt) {
    const u128 v = src[i];
    if (v != 0 && key_score_bits(v) >= theta) {
      const u32 pos = atomicAdd(&s_n, 1u);
      if (pos < static_cast<u32>(kMergeCap)) buf[pos] = v;
    }
  }
  __syncthreads();
  const u32 n = s_n;
  const u128* res = top;
  if (n <= static_cast<u32>(kMergeCap)) {
    block_topk_hist(buf, static_cast<int>(n), kMergeCap, k, hist, top, 1024);
  } else {  // heavy ties at the threshold: the streaming select of the shared helper
    block_select_lists(src, n_lists, k, k, buf, kMergeCap, ctl);
    res = buf;
  }
  for (int i = tid; i < k; i += nt) {
    const u128 key = res[i];
    const size_t o = static_cast<size_t>(qi) * k + i;
    if (key == 0) {
      S[o] = -1.7976931348623157e308;
      I[o] = -1;
    } else {
      S[o] = ordered_to_double(static_cast<u64>(key >> 64));
      I[o] = static_cast<long long>(0xFFFFFFFFu - static_cast<u32>(key)) + id_offset;
    }
  }
  if (tid == 0) tau_g[ql] = 0;
}

// block index: skip[t][b] = number of postings of term t with doc id < b * kStreamBlk
__global__ void bm25_skip_kernel(const long long* __restrict__ indptr, const int* __restrict__ doc_ids, int n_terms,
                                 int n_blocks, u32* __restrict__ skip) {
  const long long total = static_cast<long long>(n_terms) * (n_blocks + 1);
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int t = static_cast<int>(i / (n_blocks + 1));
    const int b = static_cast<int>(i - static_cast<long long>(t) * (n_blocks + 1));
    const long long ib = indptr[t], ie = indptr[t + 1];
    skip[i] = static_cast<u32>(lower_bound_i32(doc_ids, ib, ie, static_cast<long long>(b) * kStreamBlk) - ib);
  }
}

// build-time: weight[p] = tf*(k1+1) / (tf + k1*(1 - b + b*dl/avgdl)), numpy's operation order
template <class W>
__global__ void bm25_weight_kernel(const int* __restrict__ doc_ids, const int* __restrict__ tf,
                                   const int* __restrict__ doc_len, long long n_post, double k1p1, double k1,
                                   double one_minus_b, double b, double avgdl, W* __restrict__ out) {
  const long long p = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (p >= n_post) return;
  const double f = static_cast<double>(tf[p]);
  const double dl = static_cast<double>(doc_len[doc_ids[p]]);
  const double num = __dmul_rn(f, k1p1);
  const double norm = __dadd_rn(one_minus_b, __ddiv_rn(__dmul_rn(b, dl), avgdl));
  const double den = __dadd_rn(f, __dmul_rn(k1, norm));
  out[p] = static_cast<W>(__ddiv_rn(num, den));
}

struct Bm25Ws {
  unsigned char* base = nullptr;
  size_t bytes = 0;
};

}  // namespace b2r

using namespace b2r;

struct b2r_bm25 {
  int device = 0;
  long long n_docs = 0;
  int n_terms = 0;
  long long n_post = 0;
  int weight_bits = 32;
  long long* indptr = nullptr;
  int* doc_ids = nullptr;
  void* weight = nullptr;
  double* idf = nullptr;
  unsigned* skip = nullptr;   // block index [n_terms][n_blocks + 1] of the streaming kernel (may be absent)
  int n_blocks = 0;           // ceil(n_docs / kStreamBlk)
  bool owns_doc_ids = true;
  int num_sms = 148;
  std::mutex mu;
  // Host entry points are re-entrant: every call borrows a "lane" (its own stream, grow-only device + pinned
  // staging, and — through the stream-keyed workspace pool — its own kernel workspace) from a pool shared by all
  // BM25 indexes of the device, so concurrent callers of one index run concurrently (the reference calls one index
  // from a thread pool, src/retrieval.py:1498-1504), no call pays a cudaMalloc, and scratch memory follows the
  // callers' concurrency, not the number of indexes.
  struct HostLane {
    cudaStream_t stream = nullptr;
    unsigned char* dev = nullptr;
    unsigned char* pin = nullptr;
    size_t cap = 0;
  };
};

namespace b2r {

static const size_t kBm25WsBytes = static_cast<size_t>(256) << 20;  // partial lists + tickets + thresholds
static const size_t kBm25CtlBytes = 64 * 1024;                      // tickets[<=4096] + tau_g[<=4096]
static const int kBm25MaxQPerLaunch = 4096;

static std::mutex g_lane_mu;
static std::map<int, std::vector<b2r_bm25::HostLane*>> g_free_lanes;  // per device

static int bm25_lane_acquire(b2r_bm25* h, size_t bytes, b2r_bm25::HostLane** out) {
  b2r_bm25::HostLane* lane = nullptr;
  {
    std::lock_guard<std::mutex> lock(g_lane_mu);
    auto& pool = g_free_lanes[h->device];
    if (!pool.empty()) {
      lane = pool.back();
      pool.pop_back();
    }
  }
  if (!lane) {
    lane = new b2r_bm25::HostLane();
    cudaError_t e = cudaStreamCreateWithFlags(&lane->stream, cudaStreamNonBlocking);
    if (e != cudaSuccess) {
      delete lane;
      return fail(B2R_ERR_CUDA, std::string("bm25: cudaStreamCreate: ") + cudaGetErrorString(e));
    }
  }
  if (lane->cap < bytes) {
    if (lane->dev) cudaFree(lane->dev);
    if (lane->pin) cudaFreeHost(lane->pin);
    lane->dev = lane->pin = nullptr;
    lane->cap = 0;
    const size_t cap = std::max<size_t>(bytes + bytes / 4, 1 << 16);
    cudaError_t e = cudaMalloc(&lane->dev, cap);
    if (e == cudaSuccess) e = cudaMallocHost(&lane->pin, cap);
    if (e != cudaSuccess) {
      if (lane->dev) cudaFree(lane->dev);
      lane->dev = nullptr;
      std::lock_guard<std::mutex> lock(g_lane_mu);
      g_free_lanes[h->device].push_back(lane);
      return fail(e == cudaErrorMemoryAllocation ? B2R_ERR_NOMEM : B2R_ERR_CUDA, std::string("bm25: staging: ") + cudaGetErrorString(e));
    }
    lane->cap = cap;
  }
  *out = lane;
  return B2R_OK;
}

static void bm25_lane_release(b2r_bm25* h, b2r_bm25::HostLane* lane) {
  std::lock_guard<std::mutex> lock(g_lane_mu);
  g_free_lanes[h->device].push_back(lane);
}

static int bm25_get_ws(b2r_bm25* h, cudaStream_t stream, Bm25Ws* out) {
  // small corpora need far less than the cap: nq_launch * blocks * k * 16 with nq_launch <= 4096.  The buffer is the
  // stream's shared BM25 workspace (ws_pool_get): it grows to the largest index searched on that stream.
  const long long blocks = std::max<long long>(1, (h->n_docs + 8191) / 8192);
  size_t want = kBm25CtlBytes + static_cast<size_t>(std::min<long long>(blocks, 1 << 20)) * B2R_MAX_K * 16 * 4;
  want = 2 * std::min(kBm25WsBytes + kBm25CtlBytes, std::max<size_t>(want, 8 << 20));  // two halves, see bm25_run
  return ws_pool_get(h->device, stream, 1, want, want, &out->base, &out->bytes);
}

struct Bm25Plan {
  int block_docs, cap, n_sel_warps, sel_cap, grid_x;
  size_t smem;
};

static int pow2_ge_i(int v) {
  int p = 1;
  while (p < v) p <<= 1;
  return p;
}

static int bm25_plan(const b2r_bm25* h, int k, Bm25Plan* p) {
  int B = tunable("bm25.block_docs");
  if (B <= 0) B = 8192;
  B = std::max(1024, pow2_ge_i(B));
  p->cap = pow2_ge_i(std::max(2 * k, 64));
  // the accumulator area doubles as the final-select scratch: >= 2 * pow2(k) 128-bit keys
  while (static_cast<size_t>(B) * 8 < static_cast<size_t>(2) * pow2_ge_i(k) * 16) B <<= 1;
  p->n_sel_warps = kBm25Warps;
  while (p->n_sel_warps > 1 && static_cast<size_t>(p->n_sel_warps) * p->cap * 16 > 64 * 1024) p->n_sel_warps >>= 1;
  p->block_docs = B;
  p->sel_cap = 1;
  while (static_cast<size_t>(p->sel_cap) * 2 * 16 <= static_cast<size_t>(B) * 8) p->sel_cap <<= 1;
  const int cap_limit = tunable("select.buf_cap");
  if (cap_limit > 0) p->sel_cap = std::max(2 * pow2_ge_i(k), std::min(p->sel_cap, pow2_ge_i(cap_limit)));
  p->smem = static_cast<size_t>(B) * 8 + 2 * kTermChunk * 8 + 4 * 8 + static_cast<size_t>(p->n_sel_warps) * p->cap * 16 + 16;
  if (p->smem > 226 * 1024) return fail(B2R_ERR_UNSUPPORTED, "bm25: no shared-memory configuration fits");
  const long long gx = std::max<long long>(1, (h->n_docs + B - 1) / B);
  if (gx > 0x7FFFFFFFLL) return fail(B2R_ERR_UNSUPPORTED, "bm25: too many documents for one shard");
  p->grid_x = static_cast<int>(gx);
  return B2R_OK;
}

template <class W>
static int bm25_launch_t(b2r_bm25* h, const Bm25Plan& p, Bm25Args<W> a, int nq_launch, cudaStream_t stream,
                         bool overlap_prev) {
  a.weight = static_cast<const W*>(h->weight);
  auto kern = bm25_kernel<W>;
  B2R_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(p.smem)));
  B2R_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
  dim3 grid(p.grid_x, nq_launch);
  if (overlap_prev) {
    // programmatic stream serialization: this grid may start as soon as every CTA of the previous kernel in
    // the stream has executed griddepcontrol.launch_dependents (or exited).  The kernel never waits on the
    // previous grid (it reads nothing that grid writes); the next ordinary launch waits for both.
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = dim3(kBm25Warps * 32);
    cfg.dynamicSmemBytes = p.smem;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    B2R_CUDA(cudaLaunchKernelEx(&cfg, kern, a));
  } else {
    kern<<<grid, kBm25Warps * 32, p.smem, stream>>>(a);
  }
  B2R_CUDA(cudaGetLastError());
  return B2R_OK;
}

template <class W, int CHUNK, int STAGES, int MINB>
static int bm25_stream_launch_cfg(b2r_bm25* h, Bm25StreamArgs<W> a, cudaStream_t stream, bool overlap_prev) {
  a.weight = static_cast<const W*>(h->weight);
  auto kern = bm25_stream_kernel<W, CHUNK, STAGES, MINB>;
  const size_t smem = kStreamBlk * sizeof(double) + static_cast<size_t>(STAGES) * CHUNK * (4 + sizeof(W)) +
                      STAGES * sizeof(StreamMeta) + 2 * STAGES * 8 + 8 * 8;  // select scratch: inside the accumulators
  B2R_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
  B2R_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
  const long long n_items = static_cast<long long>(a.nq_launch) * a.b_count;
  int per_sm = static_cast<int>(std::max<size_t>(1, std::min<size_t>(MINB, (227 * 1024) / (smem + 1024))));
  if (tunable("bm25.ctas_per_sm") > 0) per_sm = std::min(per_sm, tunable("bm25.ctas_per_sm"));
  const int grid = static_cast<int>(std::max<long long>(1, std::min<long long>(n_items, static_cast<long long>(h->num_sms) * per_sm)));
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(kStreamThreads);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = overlap_prev ? 1 : 0;  // see bm25_launch_t
  B2R_CUDA(cudaLaunchKernelEx(&cfg, kern, a));
  B2R_CUDA(cudaGetLastError());
  return B2R_OK;
}

// Ring geometry (postings per stage, stages) and resident CTAs per SM.  Two things set the kernel's time: resident CTAs
// (it waits more than it issues: 1 / 2 / 3 CTAs per SM = 8.8 / 5.0 / 4.2 ms for 64 queries over 10M documents, r01-r02)
// and the number of stage hand-overs per item (256-posting stages x 12: +35 %, 128 x 24: +140 %).  With the select's
// scratch inside the accumulator area the ring has 40 KiB at three CTAs per SM: 1024-posting stages x 5 for float32
// weights (64 queries 2.65 -> 2.38 ms), x 3 for float64 weights (12 B / posting; three CTAs instead of two: 3.42 -> 2.63 ms).
// bm25.cfg = 1 / 2 select the earlier geometries for A/B runs.
static int bm25_stream_launch(b2r_bm25* h, Bm25StreamArgs<float> a, cudaStream_t stream, bool overlap_prev) {
  switch (tunable("bm25.cfg")) {
    case 1: return bm25_stream_launch_cfg<float, 1024, 6, 2>(h, a, stream, overlap_prev);
    case 2: return bm25_stream_launch_cfg<float, 512, 6, 3>(h, a, stream, overlap_prev);
    default: return bm25_stream_launch_cfg<float, 1024, 5, 3>(h, a, stream, overlap_prev);
  }
}
static int bm25_stream_launch(b2r_bm25* h, Bm25StreamArgs<double> a, cudaStream_t stream, bool overlap_prev) {
  switch (tunable("bm25.cfg")) {
    case 1: return bm25_stream_launch_cfg<double, 1024, 5, 2>(h, a, stream, overlap_prev);
    case 2: return bm25_stream_launch_cfg<double, 512, 6, 2>(h, a, stream, overlap_prev);
    default: return bm25_stream_launch_cfg<double, 1024, 3, 3>(h, a, stream, overlap_prev);
  }
}

static int bm25_run_stream(b2r_bm25* h, const int* term_ids_dev, const int* term_offsets_dev, int nq, int k, double* S_dev,
                           long long* I_dev, long long id_offset, double* dense_dev, cudaStream_t stream,
                           bool overlap_prev) {
  Bm25Ws ws;
  int rc = bm25_get_ws(h, stream, &ws);
  if (rc) return rc;
  // per query: the (block) lists + the k threshold slots
  const bool use_slots = !dense_dev && !tunable("bm25.no_slots");
  const size_t per_q = static_cast<size_t>(h->n_blocks) * k * 16 + (use_slots ? static_cast<size_t>(k) * 8 : 0);
  // the first half of the stream's workspace only: the second half carries the doc-block kernel's alternate control words
  long long max_q = static_cast<long long>(((ws.bytes / 2 & ~static_cast<size_t>(255)) - kBm25CtlBytes - 256) / per_q);
  max_q = std::min<long long>(max_q, kBm25MaxQPerLaunch);
  if (max_q < 1) return fail(B2R_ERR_UNSUPPORTED, "bm25: workspace too small for one query (corpus x k too large)");
  for (int q0 = 0; q0 < nq; q0 += static_cast<int>(max_q)) {
    Bm25StreamArgs<float> a;
    a.indptr = h->indptr;
    a.doc_ids = h->doc_ids;
    a.weight = nullptr;
    a.idf = h->idf;
    a.skip = tunable("bm25.no_skip") ? nullptr : h->skip;
    a.n_docs = h->n_docs;
    a.n_blocks = h->n_blocks;
    a.term_ids = term_ids_dev;
    a.term_offsets = term_offsets_dev;
    a.q_base = q0;
    a.nq_launch = static_cast<int>(std::min<long long>(max_q, nq - q0));
    a.k = k;
    a.tau_g = reinterpret_cast<u64*>(ws.base + kBm25CtlBytes / 2);
    // [slots: nq_launch x k, cleared below — the doc-block kernel keeps its own lists in this workspace][lists]
    const size_t slot_bytes = use_slots ? ((static_cast<size_t>(a.nq_launch) * k * 8 + 255) & ~static_cast<size_t>(255)) : 0;
    a.slots = use_slots ? reinterpret_cast<u64*>(ws.base + kBm25CtlBytes) : nullptr;
    a.partial = reinterpret_cast<u128*>(ws.base + kBm25CtlBytes + slot_bytes);
    a.dense = dense_dev;
    a.debug_skip = tunable("bm25.debug_skip");
    if (use_slots) B2R_CUDA(cudaMemsetAsync(a.slots, 0, static_cast<size_t>(a.nq_launch) * k * 8, stream));
    // Pilot launch.  Every query's first wave of blocks meets no threshold and pays the full histogram select; with
    // many queries in a launch that is the whole grid once per query.  So the first few blocks of EVERY query go first
    // (about one wave in all): when the main launch reaches a query, its slots already hold a threshold from those
    // blocks and its first wave selects among a handful of documents instead of 4096.  One query alone gains nothing
    // (its first wave is a fifth of its blocks either way), so small batches keep the single launch.
    int pilot = 0;
    if (use_slots && a.nq_launch >= 4) {
      pilot = tunable("bm25.pilot_blocks");
      if (pilot == 0) pilot = std::max(1, 3 * h->num_sms / a.nq_launch);
      if (pilot < 0 || pilot * 4 > a.n_blocks) pilot = 0;
    }
    for (int part = pilot > 0 ? 0 : 1; part < 2; ++part) {
      a.b_begin = part == 0 ? 0 : pilot;
      a.b_count = part == 0 ? pilot : a.n_blocks - pilot;
      if (h->weight_bits == 64) {
        Bm25StreamArgs<double> ad;
        memcpy(&ad, &a, sizeof(a));  // identical layout: only the weight pointer type differs
        rc = bm25_stream_launch(h, ad, stream, overlap_prev && q0 == 0 && part == (pilot > 0 ? 0 : 1));
      } else {
        rc = bm25_stream_launch(h, a, stream, overlap_prev && q0 == 0 && part == (pilot > 0 ? 0 : 1));
      }
      if (rc) return rc;
      ++launch_counter();
    }
    if (!dense_dev) {
      const int msmem = kMergeCap * 16 + 1024 * 16 + (kSelBins + 4) * 4;
      auto mk = bm25_merge_kernel;
      B2R_CUDA(cudaFuncSetAttribute(mk, cudaFuncAttributeMaxDynamicSharedMemorySize, msmem));
      mk<<<a.nq_launch, 1024, msmem, stream>>>(a.partial, a.n_blocks, k, q0, a.tau_g, S_dev, I_dev, id_offset);
      B2R_CUDA(cudaGetLastError());
      ++launch_counter();
    }
  }
  return B2R_OK;
}

// shared by topk (dense == nullptr) and scores (dense != nullptr, one query)
// Consecutive doc-block launches on one stream alternate between the two halves of the stream's workspace: with
// chained steps (HybridTail::trigger_next) the BM25 kernel of step s + 1 may start while the one of step s has not
// finished its tail, and must not share partial lists or control words with it.
static std::mutex g_parity_mu;
static std::map<cudaStream_t, int> g_ws_parity;

static int bm25_run(b2r_bm25* h, const int* term_ids_dev, const int* term_offsets_dev, int nq, int k, double* S_dev,
                    long long* I_dev, long long id_offset, double* dense_dev, cudaStream_t stream,
                    bool overlap_prev = false, const HybridTail* tail = nullptr) {
  launch_counter() = 0;
  // Alone, the streaming kernel is ~1.6x faster (profiles/r01_bm25.md); launched BESIDE the flat scan its busy
  // consumer warps take more from the scan than the latency-bound doc-block kernel does, so the hybrid
  // single-query path keeps the latter (bm25.v1 = 2 forces the streaming kernel there too).
  const int v1 = tunable("bm25.v1");
  const bool use_stream = v1 == 2 || (v1 == 0 && !(overlap_prev && nq <= 4));
  if (tail && (use_stream || k > kTailMaxK || dense_dev)) {
    return fail(B2R_ERR_UNSUPPORTED, "bm25: the fused tail needs the doc-block kernel (<= 4 queries beside a scan) and k <= 256");
  }
  if (use_stream && k <= kStreamMaxK) {
    return bm25_run_stream(h, term_ids_dev, term_offsets_dev, nq, k, S_dev, I_dev, id_offset, dense_dev, stream, overlap_prev);
  }
  Bm25Plan p;
  int rc = bm25_plan(h, k, &p);
  if (rc) return rc;
  Bm25Ws ws;
  rc = bm25_get_ws(h, stream, &ws);
  if (rc) return rc;
  const size_t per_q = static_cast<size_t>(p.grid_x) * k * 16;
  const size_t half = (ws.bytes / 2) & ~static_cast<size_t>(255);
  long long max_q = static_cast<long long>((half - kBm25CtlBytes) / per_q);
  max_q = std::min<long long>(std::min<long long>(max_q, kBm25MaxQPerLaunch), 65535);
  if (max_q < 1) return fail(B2R_ERR_UNSUPPORTED, "bm25: workspace too small for one query (corpus x k too large)");
  if (tail && nq > max_q) return fail(B2R_ERR_UNSUPPORTED, "bm25: the fused tail needs all queries in one launch");
  int parity;
  {
    std::lock_guard<std::mutex> lock(g_parity_mu);
    parity = (g_ws_parity[stream] ^= 1);
  }
  unsigned char* ws_base = ws.base + static_cast<size_t>(parity) * half;
  for (int q0 = 0; q0 < nq; q0 += static_cast<int>(max_q)) {
    const int nql = static_cast<int>(std::min<long long>(max_q, nq - q0));
    Bm25Args<float> a;
    a.indptr = h->indptr;
    a.doc_ids = h->doc_ids;
    a.weight = nullptr;
    a.idf = h->idf;
    a.n_docs = h->n_docs;
    a.block_docs = p.block_docs;
    a.term_ids = term_ids_dev;
    a.term_offsets = term_offsets_dev;
    a.q_base = q0;
    a.k = k;
    a.cap = p.cap;
    a.n_sel_warps = p.n_sel_warps;
    a.tickets = reinterpret_cast<u32*>(ws_base);
    a.tau_g = reinterpret_cast<u64*>(ws_base + kBm25CtlBytes / 2);
    a.partial = reinterpret_cast<u128*>(ws_base + kBm25CtlBytes);
    a.S = S_dev;
    a.I = I_dev;
    a.id_offset = id_offset;
    a.dense = dense_dev;
    a.sel_cap = p.sel_cap;
    if (tail) a.tail = *tail;
    else memset(&a.tail, 0, sizeof(a.tail));
    if (h->weight_bits == 64) {
      Bm25Args<double> ad;
      memcpy(&ad, &a, sizeof(a));  // identical layout: only the weight pointer type differs
      rc = bm25_launch_t<double>(h, p, ad, nql, stream, overlap_prev && q0 == 0);
    } else {
      rc = bm25_launch_t<float>(h, p, a, nql, stream, overlap_prev && q0 == 0);
    }
    if (rc) return rc;
    ++launch_counter();
  }
  return B2R_OK;
}

}  // namespace b2r

int bm25_topk_with_tail(b2r_bm25* h, const int32_t* term_ids_dev, const int32_t* term_offsets_dev, int nq, int k,
                        double* S_dev, int64_t* J_dev, int64_t id_offset, cudaStream_t stream, bool overlap_prev,
                        const b2r::HybridTail& tail) {
  if (!h) return b2r::fail(B2R_ERR_INVALID, "bm25: NULL handle");
  return b2r::bm25_run(h, term_ids_dev, term_offsets_dev, nq, k, S_dev, reinterpret_cast<long long*>(J_dev), id_offset, nullptr,
                       stream, overlap_prev, &tail);
}

extern "C" {

// weight_bits and device-resident inputs are selected through tunables-free explicit arguments in
// b2r_bm25_create_ex; b2r_bm25_create is the plain host-input, float64-weight (rank_bm25-exact) form of the header.
int b2r_bm25_create_ex(int64_t n_docs, int32_t n_terms, const int64_t* indptr, const int32_t* doc_ids,
                       const int32_t* tf, const int32_t* doc_len, double avgdl, const double* idf, double k1, double b,
                       int device, int inputs_on_device, int weight_bits, b2r_bm25** out) {
  if (!out) return fail(B2R_ERR_INVALID, "b2r_bm25_create: out is NULL");
  *out = nullptr;
  if (n_docs < 1 || n_terms < 1 || !indptr || !doc_len || !idf) return fail(B2R_ERR_INVALID, "b2r_bm25_create: bad arguments");
  if (n_docs > 0xFFFFFFF0LL) return fail(B2R_ERR_UNSUPPORTED, "b2r_bm25_create: more than 2^32-16 documents in one shard");
  if (weight_bits != 32 && weight_bits != 64) return fail(B2R_ERR_INVALID, "b2r_bm25_create: weight_bits must be 32 or 64");
  if (!(avgdl > 0.0)) return fail(B2R_ERR_INVALID, "b2r_bm25_create: avgdl must be positive");
  int ndev = 0;
  B2R_CUDA(cudaGetDeviceCount(&ndev));
  if (device < 0 || device >= ndev) return fail(B2R_ERR_INVALID, "b2r_bm25_create: no such CUDA device");
  DeviceGuard g(device);
  if (!g.ok) return fail(B2R_ERR_CUDA, "b2r_bm25_create: cudaSetDevice failed");
  cudaDeviceProp prop;
  B2R_CUDA(cudaGetDeviceProperties(&prop, device));
  if (prop.major != 10) return fail(B2R_ERR_UNSUPPORTED, "libb2r is built for sm_100a (B200) only");

  const cudaMemcpyKind kind = inputs_on_device ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice;
  long long n_post = 0;
  if (inputs_on_device) {
    B2R_CUDA(cudaMemcpy(&n_post, indptr + n_terms, sizeof(long long), cudaMemcpyDeviceToHost));
  } else {
    n_post = indptr[n_terms];
  }
  if (n_post < 0 || (n_post > 0 && (!doc_ids || !tf))) return fail(B2R_ERR_INVALID, "b2r_bm25_create: bad postings");

  b2r_bm25* h = new b2r_bm25();
  h->device = device;
  h->n_docs = n_docs;
  h->n_terms = n_terms;
  h->n_post = n_post;
  h->weight_bits = weight_bits;
  h->num_sms = prop.multiProcessorCount;
  int* tf_dev = nullptr;
  int* dl_dev = nullptr;
  const size_t np_alloc = static_cast<size_t>(std::max<long long>(n_post, 1)) + 4;
  cudaError_t e = cudaMalloc(&h->indptr, (static_cast<size_t>(n_terms) + 1) * sizeof(long long));
  if (e == cudaSuccess) e = cudaMalloc(&h->idf, static_cast<size_t>(n_terms) * sizeof(double));
  if (e == cudaSuccess) e = cudaMalloc(&h->weight, np_alloc * (weight_bits / 8));
  if (e == cudaSuccess) e = cudaMemcpy(h->indptr, indptr, (static_cast<size_t>(n_terms) + 1) * sizeof(long long), kind);
  if (e == cudaSuccess) e = cudaMemcpy(h->idf, idf, static_cast<size_t>(n_terms) * sizeof(double), kind);
  if (e == cudaSuccess) {
    if (inputs_on_device) {
      // borrow the caller's device arrays for the build; keep a private copy of doc_ids
      e = cudaMalloc(&h->doc_ids, np_alloc * sizeof(int));
      if (e == cudaSuccess && n_post) e = cudaMemcpy(h->doc_ids, doc_ids, static_cast<size_t>(n_post) * sizeof(int), kind);
      tf_dev = const_cast<int*>(tf);
      dl_dev = const_cast<int*>(doc_len);
    } else {
      e = cudaMalloc(&h->doc_ids, np_alloc * sizeof(int));
      if (e == cudaSuccess) e = cudaMalloc(&tf_dev, np_alloc * sizeof(int));
      if (e == cudaSuccess) e = cudaMalloc(&dl_dev, static_cast<size_t>(n_docs) * sizeof(int));
      if (e == cudaSuccess && n_post) e = cudaMemcpy(h->doc_ids, doc_ids, static_cast<size_t>(n_post) * sizeof(int), kind);
      if (e == cudaSuccess && n_post) e = cudaMemcpy(tf_dev, tf, static_cast<size_t>(n_post) * sizeof(int), kind);
      if (e == cudaSuccess) e = cudaMemcpy(dl_dev, doc_len, static_cast<size_t>(n_docs) * sizeof(int), kind);
    }
  }
  if (e == cudaSuccess && n_post) {
    const unsigned blocks = static_cast<unsigned>((n_post + 255) / 256);
    if (weight_bits == 64) {
      bm25_weight_kernel<double><<<blocks, 256>>>(h->doc_ids, tf_dev, dl_dev, n_post, k1 + 1.0, k1, 1.0 - b, b, avgdl,
                                                  static_cast<double*>(h->weight));
    } else {
      bm25_weight_kernel<float><<<blocks, 256>>>(h->doc_ids, tf_dev, dl_dev, n_post, k1 + 1.0, k1, 1.0 - b, b, avgdl,
                                                 static_cast<float*>(h->weight));
    }
    e = cudaGetLastError();
    if (e == cudaSuccess) e = cudaDeviceSynchronize();
  }
  if (!inputs_on_device) {
    if (tf_dev) cudaFree(tf_dev);
    if (dl_dev) cudaFree(dl_dev);
  }
  if (e == cudaSuccess) {
    // block index of the streaming kernel: 4 bytes per (term, 4096-document block); skipped when it would
    // outweigh a quarter of the postings (huge vocabularies over huge shards fall back to binary searches)
    h->n_blocks = static_cast<int>((n_docs + kStreamBlk - 1) / kStreamBlk);
    const size_t skip_bytes = static_cast<size_t>(n_terms) * (static_cast<size_t>(h->n_blocks) + 1) * sizeof(unsigned);
    const size_t post_bytes = static_cast<size_t>(n_post) * (4 + weight_bits / 8);
    if (skip_bytes <= std::max<size_t>(static_cast<size_t>(256) << 20, post_bytes / 4)) {
      e = cudaMalloc(&h->skip, skip_bytes);
      if (e == cudaSuccess) {
        bm25_skip_kernel<<<prop.multiProcessorCount * 8, 256>>>(h->indptr, h->doc_ids, n_terms, h->n_blocks, h->skip);
        e = cudaGetLastError();
        if (e == cudaSuccess) e = cudaDeviceSynchronize();
      }
    }
  }
  if (e != cudaSuccess) {
    const int code = e == cudaErrorMemoryAllocation ? B2R_ERR_NOMEM : B2R_ERR_CUDA;
    std::string msg = std::string("b2r_bm25_create: ") + cudaGetErrorString(e);
    b2r_bm25_destroy(h);
    return fail(code, msg);
  }
  *out = h;
  return B2R_OK;
}

int b2r_bm25_create(int64_t n_docs, int32_t n_terms, const int64_t* indptr, const int32_t* doc_ids, const int32_t* tf,
                    const int32_t* doc_len, double avgdl, const double* idf, double k1, double b, int device,
                    b2r_bm25** out) {
  return b2r_bm25_create_ex(n_docs, n_terms, indptr, doc_ids, tf, doc_len, avgdl, idf, k1, b, device, 0, 64, out);
}

int b2r_bm25_ndocs(const b2r_bm25* h, int64_t* out) {
  if (!h || !out) return fail(B2R_ERR_INVALID, "b2r_bm25_ndocs: NULL argument");
  *out = h->n_docs;
  return B2R_OK;
}

int b2r_bm25_postings(const b2r_bm25* h, int64_t* n_post, int* weight_bits) {
  if (!h) return fail(B2R_ERR_INVALID, "b2r_bm25_postings: NULL handle");
  if (n_post) *n_post = h->n_post;
  if (weight_bits) *weight_bits = h->weight_bits;
  return B2R_OK;
}

int bm25_check_terms(const b2r_bm25* h, const int32_t* term_ids, int n) {
  for (int i = 0; i < n; ++i) {
    if (term_ids[i] < 0 || term_ids[i] >= h->n_terms) {
      return fail(B2R_ERR_INVALID, "bm25: term id " + std::to_string(term_ids[i]) + " out of range");
    }
  }
  return B2R_OK;
}

int b2r_bm25_topk_dev(b2r_bm25* h, const int32_t* term_ids_dev, const int32_t* term_offsets_dev,
                      const int32_t* term_offsets_host, int nq, int k, double* S_dev, int64_t* I_dev, int64_t id_offset,
                      void* stream) {
  return b2r_bm25_topk_dev_ex(h, term_ids_dev, term_offsets_dev, term_offsets_host, nq, k, S_dev, I_dev, id_offset, stream,
                              0u);
}

int b2r_bm25_topk_dev_ex(b2r_bm25* h, const int32_t* term_ids_dev, const int32_t* term_offsets_dev,
                         const int32_t* term_offsets_host, int nq, int k, double* S_dev, int64_t* I_dev,
                         int64_t id_offset, void* stream, unsigned flags) {
  (void)term_offsets_host;
  if (!h) return fail(B2R_ERR_INVALID, "b2r_bm25_topk: NULL handle");
  if (nq < 0 || k < 1) return fail(B2R_ERR_INVALID, "b2r_bm25_topk: need nq >= 0 and k >= 1");
  if (k > B2R_MAX_K) return fail(B2R_ERR_UNSUPPORTED, "b2r_bm25_topk: k > B2R_MAX_K (1024)");
  if (nq == 0) return B2R_OK;
  if (!term_offsets_dev || !S_dev || !I_dev) return fail(B2R_ERR_INVALID, "b2r_bm25_topk: NULL buffer");
  DeviceGuard g(h->device);
  if (!g.ok) return fail(B2R_ERR_CUDA, "b2r_bm25_topk: cudaSetDevice failed");
  return bm25_run(h, term_ids_dev, term_offsets_dev, nq, k, S_dev, reinterpret_cast<long long*>(I_dev), id_offset,
                  nullptr, static_cast<cudaStream_t>(stream), (flags & B2R_LAUNCH_OVERLAP_PREVIOUS) != 0u);
}

int b2r_bm25_topk_host(b2r_bm25* h, const int32_t* term_ids, const int32_t* term_offsets, int nq, int k, double* S,
                       int64_t* I) {
  if (!h) return fail(B2R_ERR_INVALID, "b2r_bm25_topk: NULL handle");
  if (nq < 0 || k < 1) return fail(B2R_ERR_INVALID, "b2r_bm25_topk: need nq >= 0 and k >= 1");
  if (k > B2R_MAX_K) return fail(B2R_ERR_UNSUPPORTED, "b2r_bm25_topk: k > B2R_MAX_K (1024)");
  if (nq == 0) return B2R_OK;
  if (!term_offsets || !S || !I) return fail(B2R_ERR_INVALID, "b2r_bm25_topk: NULL buffer");
  const int n_terms_total = term_offsets[nq];
  if (term_offsets[0] != 0 || n_terms_total < 0 || (n_terms_total > 0 && !term_ids)) return fail(B2R_ERR_INVALID, "b2r_bm25_topk: bad term offsets");
  int rc = bm25_check_terms(h, term_ids, n_terms_total);
  if (rc) return rc;
  DeviceGuard g(h->device);
  if (!g.ok) return fail(B2R_ERR_CUDA, "b2r_bm25_topk: cudaSetDevice failed");
  const size_t nk = static_cast<size_t>(nq) * k;
  // staging layout (device and pinned alike): [S f64 nk][I i64 nk][offsets i32 nq+1][term ids i32]
  const size_t off_o = nk * 16, off_t = off_o + ((static_cast<size_t>(nq) + 1) * 4 + 15) / 16 * 16;
  const size_t in_bytes = off_t + static_cast<size_t>(std::max(n_terms_total, 1)) * 4 - off_o;
  b2r_bm25::HostLane* lane = nullptr;
  rc = bm25_lane_acquire(h, off_o + in_bytes, &lane);
  if (rc) return rc;
  cudaStream_t st = lane->stream;
  memcpy(lane->pin + off_o, term_offsets, (static_cast<size_t>(nq) + 1) * 4);
  if (n_terms_total) memcpy(lane->pin + off_t, term_ids, static_cast<size_t>(n_terms_total) * 4);
  cudaError_t e = cudaMemcpyAsync(lane->dev + off_o, lane->pin + off_o, in_bytes, cudaMemcpyHostToDevice, st);
  if (e == cudaSuccess) {
    rc = bm25_run(h, reinterpret_cast<const int*>(lane->dev + off_t), reinterpret_cast<const int*>(lane->dev + off_o), nq, k,
                  reinterpret_cast<double*>(lane->dev), reinterpret_cast<long long*>(lane->dev + nk * 8), 0, nullptr, st);
    if (rc == B2R_OK) {
      e = cudaMemcpyAsync(lane->pin, lane->dev, nk * 16, cudaMemcpyDeviceToHost, st);
      if (e == cudaSuccess) e = cudaStreamSynchronize(st);
      if (e == cudaSuccess) {
        memcpy(S, lane->pin, nk * 8);
        memcpy(I, lane->pin + nk * 8, nk * 8);
      }
    }
  }
  if (e != cudaSuccess) cudaStreamSynchronize(st);  // nothing may still be using the lane when it goes back
  bm25_lane_release(h, lane);
  if (rc == B2R_OK && e != cudaSuccess) rc = fail(B2R_ERR_CUDA, std::string("b2r_bm25_topk_host: ") + cudaGetErrorString(e));
  return rc;
}

int b2r_bm25_scores_host(b2r_bm25* h, const int32_t* term_ids, int n_q_terms, double* scores) {
  if (!h) return fail(B2R_ERR_INVALID, "b2r_bm25_scores: NULL handle");
  if (n_q_terms < 0 || !scores || (n_q_terms > 0 && !term_ids)) return fail(B2R_ERR_INVALID, "b2r_bm25_scores: bad arguments");
  int rc = bm25_check_terms(h, term_ids, n_q_terms);
  if (rc) return rc;
  DeviceGuard g(h->device);
  if (!g.ok) return fail(B2R_ERR_CUDA, "b2r_bm25_scores: cudaSetDevice failed");
  // staging: [dense f64 n_docs][offsets i32 x 2 (16 B)][term ids]; the scores come back through the pinned half
  const size_t off_o = static_cast<size_t>(h->n_docs) * 8, off_t = off_o + 16;
  const size_t in_bytes = 16 + static_cast<size_t>(std::max(n_q_terms, 1)) * 4;
  b2r_bm25::HostLane* lane = nullptr;
  rc = bm25_lane_acquire(h, off_o + in_bytes, &lane);
  if (rc) return rc;
  cudaStream_t st = lane->stream;
  const int offs[2] = {0, n_q_terms};
  memcpy(lane->pin + off_o, offs, sizeof(offs));
  if (n_q_terms) memcpy(lane->pin + off_t, term_ids, static_cast<size_t>(n_q_terms) * 4);
  cudaError_t e = cudaMemcpyAsync(lane->dev + off_o, lane->pin + off_o, in_bytes, cudaMemcpyHostToDevice, st);
  if (e == cudaSuccess) {
    rc = bm25_run(h, reinterpret_cast<const int*>(lane->dev + off_t), reinterpret_cast<const int*>(lane->dev + off_o), 1, 1,
                  nullptr, nullptr, 0, reinterpret_cast<double*>(lane->dev), st);
    if (rc == B2R_OK) {
      e = cudaMemcpyAsync(lane->pin, lane->dev, off_o, cudaMemcpyDeviceToHost, st);
      if (e == cudaSuccess) e = cudaStreamSynchronize(st);
      if (e == cudaSuccess) memcpy(scores, lane->pin, off_o);
    }
  }
  if (e != cudaSuccess) cudaStreamSynchronize(st);
  bm25_lane_release(h, lane);
  if (rc == B2R_OK && e != cudaSuccess) rc = fail(B2R_ERR_CUDA, std::string("b2r_bm25_scores_host: ") + cudaGetErrorString(e));
  return rc;
}

int b2r_bm25_destroy(b2r_bm25* h) {
  if (!h) return B2R_OK;
  DeviceGuard g(h->device);
  cudaDeviceSynchronize();
  if (h->indptr) cudaFree(h->indptr);
  if (h->doc_ids) cudaFree(h->doc_ids);
  if (h->weight) cudaFree(h->weight);
  if (h->idf) cudaFree(h->idf);
  if (h->skip) cudaFree(h->skip);
  // host lanes and kernel workspaces belong to the device-wide pools and outlive the handle
  delete h;
  return B2R_OK;
}

}  // extern "C"
