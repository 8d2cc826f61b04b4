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

__global__ void __launch_bounds__(256) hybrid_fuse_kernel(const float* __restrict__ D, const long long* __restrict__ I,
                                                          const double* __restrict__ S, const long long* __restrict__ J,
                                                          int k_dense, int k_bm25, int k_out, float* __restrict__ F,
                                                          long long* __restrict__ O, unsigned char* __restrict__ M,
                                                          int buf_cap) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  u64* buf = reinterpret_cast<u64*>(smem_raw);          // [buf_cap] keys
  double* red = reinterpret_cast<double*>(buf + buf_cap);  // [2] min / max
  const int q = blockIdx.x, tid = threadIdx.x, nt = blockDim.x;
  const float* Dq = D + static_cast<size_t>(q) * k_dense;
  const long long* Iq = I + static_cast<size_t>(q) * k_dense;
  const double* Sq = S + static_cast<size_t>(q) * k_bm25;
  const long long* Jq = J + static_cast<size_t>(q) * k_bm25;
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
  // key low word: (0xFFFFFFFF - id); bit 0/1 method flags travel in a side table indexed by slot
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
    const size_t o = static_cast<size_t>(q) * k_out + i;
    if (key == 0) {
      F[o] = -3.4028235e38f;
      O[o] = -1;
      if (M) M[o] = 0;
    } else {
      const long long id = static_cast<long long>(key_id(key));
      F[o] = key_score(key);
      O[o] = id;
      if (M) {
        unsigned char m = 0;
        for (int t = 0; t < k_dense; ++t) m |= (Iq[t] == id) ? 1 : 0;
        for (int j = 0; j < k_bm25; ++j) m |= (Jq[j] == id) ? 2 : 0;
        M[o] = m;  // bit 0: dense hit, bit 1: BM25 hit  (-> retrieval_sources of the record)
      }
    }
  }
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
