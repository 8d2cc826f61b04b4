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
// operation order numpy uses, stored as float32 (8 B / posting, the default) or float64
// (12 B / posting, scores then equal rank_bm25's bit for bit); idf float64[V].
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
};

template <class W>
__global__ void __launch_bounds__(kBm25Warps * 32) bm25_kernel(const Bm25Args<W> a) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
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
  bool owns_doc_ids = true;
  int num_sms = 148;
  std::mutex mu;
  std::map<cudaStream_t, Bm25Ws> stream_ws;
  cudaStream_t host_stream = nullptr;
  std::mutex host_mu;  // host entry points are serialised per handle
};

namespace b2r {

static const size_t kBm25WsBytes = static_cast<size_t>(256) << 20;  // partial lists + tickets + thresholds
static const size_t kBm25CtlBytes = 64 * 1024;                      // tickets[<=4096] + tau_g[<=4096]
static const int kBm25MaxQPerLaunch = 4096;

static int bm25_get_ws(b2r_bm25* h, cudaStream_t stream, Bm25Ws* out) {
  std::lock_guard<std::mutex> lock(h->mu);
  auto it = h->stream_ws.find(stream);
  if (it != h->stream_ws.end()) {
    *out = it->second;
    return B2R_OK;
  }
  Bm25Ws ws;
  // small corpora need far less than the cap: nq_launch * blocks * k * 16 with nq_launch <= 4096
  const long long blocks = std::max<long long>(1, (h->n_docs + 8191) / 8192);
  size_t want = kBm25CtlBytes + static_cast<size_t>(std::min<long long>(blocks, 1 << 20)) * B2R_MAX_K * 16 * 4;
  ws.bytes = std::min(kBm25WsBytes + kBm25CtlBytes, std::max<size_t>(want, 8 << 20));
  B2R_CUDA(cudaMalloc(&ws.base, ws.bytes));
  B2R_CUDA(cudaMemsetAsync(ws.base, 0, kBm25CtlBytes, stream));
  h->stream_ws[stream] = ws;
  *out = ws;
  return B2R_OK;
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

// shared by topk (dense == nullptr) and scores (dense != nullptr, one query)
static int bm25_run(b2r_bm25* h, const int* term_ids_dev, const int* term_offsets_dev, int nq, int k, double* S_dev,
                    long long* I_dev, long long id_offset, double* dense_dev, cudaStream_t stream,
                    bool overlap_prev = false) {
  launch_counter() = 0;
  Bm25Plan p;
  int rc = bm25_plan(h, k, &p);
  if (rc) return rc;
  Bm25Ws ws;
  rc = bm25_get_ws(h, stream, &ws);
  if (rc) return rc;
  const size_t per_q = static_cast<size_t>(p.grid_x) * k * 16;
  long long max_q = static_cast<long long>((ws.bytes - kBm25CtlBytes) / per_q);
  max_q = std::min<long long>(std::min<long long>(max_q, kBm25MaxQPerLaunch), 65535);
  if (max_q < 1) return fail(B2R_ERR_UNSUPPORTED, "bm25: workspace too small for one query (corpus x k too large)");
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
    a.tickets = reinterpret_cast<u32*>(ws.base);
    a.tau_g = reinterpret_cast<u64*>(ws.base + kBm25CtlBytes / 2);
    a.partial = reinterpret_cast<u128*>(ws.base + kBm25CtlBytes);
    a.S = S_dev;
    a.I = I_dev;
    a.id_offset = id_offset;
    a.dense = dense_dev;
    a.sel_cap = p.sel_cap;
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

extern "C" {

// weight_bits and device-resident inputs are selected through tunables-free explicit arguments in
// b2r_bm25_create_ex; b2r_bm25_create is the plain host-input, float32-weight form of the header.
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
  const size_t np_alloc = static_cast<size_t>(std::max<long long>(n_post, 1));
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
  if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&h->host_stream, cudaStreamNonBlocking);
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
  return b2r_bm25_create_ex(n_docs, n_terms, indptr, doc_ids, tf, doc_len, avgdl, idf, k1, b, device, 0, 32, out);
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

static int bm25_check_terms(const b2r_bm25* h, const int32_t* term_ids, int n) {
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
  std::lock_guard<std::mutex> lock(h->host_mu);
  const size_t nk = static_cast<size_t>(nq) * k;
  int *t_dev = nullptr, *o_dev = nullptr;
  double* S_dev = nullptr;
  long long* I_dev = nullptr;
  cudaStream_t st = h->host_stream;
  cudaError_t e = cudaMalloc(&t_dev, std::max<size_t>(1, n_terms_total) * sizeof(int));
  if (e == cudaSuccess) e = cudaMalloc(&o_dev, (static_cast<size_t>(nq) + 1) * sizeof(int));
  if (e == cudaSuccess) e = cudaMalloc(&S_dev, nk * sizeof(double));
  if (e == cudaSuccess) e = cudaMalloc(&I_dev, nk * sizeof(long long));
  if (e == cudaSuccess && n_terms_total) e = cudaMemcpyAsync(t_dev, term_ids, static_cast<size_t>(n_terms_total) * sizeof(int), cudaMemcpyHostToDevice, st);
  if (e == cudaSuccess) e = cudaMemcpyAsync(o_dev, term_offsets, (static_cast<size_t>(nq) + 1) * sizeof(int), cudaMemcpyHostToDevice, st);
  if (e == cudaSuccess) {
    rc = bm25_run(h, t_dev, o_dev, nq, k, S_dev, I_dev, 0, nullptr, st);
    if (rc == B2R_OK) {
      e = cudaMemcpyAsync(S, S_dev, nk * sizeof(double), cudaMemcpyDeviceToHost, st);
      if (e == cudaSuccess) e = cudaMemcpyAsync(I, I_dev, nk * sizeof(long long), cudaMemcpyDeviceToHost, st);
      if (e == cudaSuccess) e = cudaStreamSynchronize(st);
    }
  }
  cudaFree(t_dev);
  cudaFree(o_dev);
  cudaFree(S_dev);
  cudaFree(I_dev);
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
  std::lock_guard<std::mutex> lock(h->host_mu);
  int *t_dev = nullptr, *o_dev = nullptr;
  double* dense = nullptr;
  cudaStream_t st = h->host_stream;
  const int offs[2] = {0, n_q_terms};
  cudaError_t e = cudaMalloc(&t_dev, std::max<size_t>(1, n_q_terms) * sizeof(int));
  if (e == cudaSuccess) e = cudaMalloc(&o_dev, 2 * sizeof(int));
  if (e == cudaSuccess) e = cudaMalloc(&dense, static_cast<size_t>(h->n_docs) * sizeof(double));
  if (e == cudaSuccess && n_q_terms) e = cudaMemcpyAsync(t_dev, term_ids, static_cast<size_t>(n_q_terms) * sizeof(int), cudaMemcpyHostToDevice, st);
  if (e == cudaSuccess) e = cudaMemcpyAsync(o_dev, offs, sizeof(offs), cudaMemcpyHostToDevice, st);
  if (e == cudaSuccess) {
    rc = bm25_run(h, t_dev, o_dev, 1, 1, nullptr, nullptr, 0, dense, st);
    if (rc == B2R_OK) {
      e = cudaMemcpyAsync(scores, dense, static_cast<size_t>(h->n_docs) * sizeof(double), cudaMemcpyDeviceToHost, st);
      if (e == cudaSuccess) e = cudaStreamSynchronize(st);
    }
  }
  cudaFree(t_dev);
  cudaFree(o_dev);
  cudaFree(dense);
  if (rc == B2R_OK && e != cudaSuccess) rc = fail(B2R_ERR_CUDA, std::string("b2r_bm25_scores_host: ") + cudaGetErrorString(e));
  return rc;
}

int b2r_bm25_destroy(b2r_bm25* h) {
  if (!h) return B2R_OK;
  DeviceGuard g(h->device);
  cudaDeviceSynchronize();
  for (auto& kv : h->stream_ws) {
    if (kv.second.base) cudaFree(kv.second.base);
  }
  if (h->indptr) cudaFree(h->indptr);
  if (h->doc_ids) cudaFree(h->doc_ids);
  if (h->weight) cudaFree(h->weight);
  if (h->idf) cudaFree(h->idf);
  if (h->host_stream) cudaStreamDestroy(h->host_stream);
  delete h;
  return B2R_OK;
}

}  // extern "C"
