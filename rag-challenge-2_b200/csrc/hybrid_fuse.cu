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

#include "hybrid_fuse.cuh"

namespace b2r {

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

__global__ void __launch_bounds__(256) hybrid_exchange_fuse_kernel(const float* __restrict__ D,
                                                                   const long long* __restrict__ I,
                                                                   const double* __restrict__ S,
                                                                   const long long* __restrict__ J, int nq, int k,
                                                                   const ExchangeArgs x, float* __restrict__ F,
                                                                   long long* __restrict__ O) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  // the next step's scan (B2R_LAUNCH_OVERLAP_PREVIOUS) may start now: this kernel's wait for the slowest rank then
  // costs nothing, the fast ranks are already streaming their next query
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  const size_t q = blockIdx.x;
  exchange_merge_fuse_query(D + q * k, I + q * k, S + q * k, J + q * k, static_cast<int>(q), k, x, F + q * k, O + q * k, smem_raw);
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
  const size_t smem = exchange_smem_bytes(world, k);
  B2R_CUDA(cudaFuncSetAttribute(hybrid_exchange_fuse_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024));
  hybrid_exchange_fuse_kernel<<<nq, 256, smem, static_cast<cudaStream_t>(stream)>>>(
      D, reinterpret_cast<const long long*>(I), S, reinterpret_cast<const long long*>(J), nq, k, x, F,
      reinterpret_cast<long long*>(O));
  B2R_CUDA(cudaGetLastError());
  return B2R_OK;
}
