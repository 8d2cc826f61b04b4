// index_build.cu — building on the GPU what the reference builds in Python, with kernels of this library:
//
//  (1) BM25 postings from a tokenised corpus — the role of BM25Okapi([chunk.split() ...]) in src/ingestion.py:19-22
//      (rank_bm25 fills one {token: tf} dictionary per document and a document-frequency dictionary):
//        token_ids[total], doc_offsets[n_docs + 1]  ->  term-major CSR (indptr, doc ids ascending per term, tf), df,
//        doc_len, and the global position of every term's first occurrence (rank_bm25 sums the idf values in
//        dictionary = first-seen order, so the average needs it).
//      Three kernels, no sort of the corpus:
//        bm25_doc_pairs_kernel   one warp per document: bitonic sort of (term, position) in shared memory, run lengths
//                                -> the document's unique (term, tf) pairs, written over its own token slots; df / first
//                                position by atomics; the count of every (term, 4096-document block) cell by atomics;
//        bm25_block_scan_kernel  one warp per term: exclusive scan of its block counts -> the term's offset of every
//                                block (this is also the block index the streaming scoring kernel uses);
//        bm25_place_kernel       one CTA per 4096-document block walks its documents IN ORDER; a pair's slot is
//                                indptr[term] + block offset + (pairs of that term already placed in this block) — the
//                                running count is an atomic counter that only this CTA touches, and because the documents
//                                are visited serially the ranks come out in document order: ascending ids, no sort.
//      df -> indptr is a single-CTA exclusive scan (the vocabulary is small next to the corpus).
//
//  (2) the down-sampled heat map of the N x N cosine matrix (generate_similarity_matrix.py:124-160 plots the full matrix):
//        heatmap_inv_norm_kernel  1 / (|x| + 1e-8) per row (fp32, as the script normalises);
//        heatmap_pool_kernel      one CTA per pixel group: float64 sum of its normalised rows (deterministic order) and of
//                                 their squared norms (for the diagonal the script forces to 1);
//        heatmap_gram_kernel      pixel (a, b) = <pool_a, pool_b> (+ the diagonal correction) / (size_a size_b), clipped.
#include <algorithm>

#include "b2r_common.cuh"

namespace b2r {

static const int kBuildBlk = 4096;        // documents per block (== the scoring kernels' block index granularity)
static const int kDocWarps = 8;           // documents per CTA of the pairs kernel
static const int kDocMaxLen = 1024;       // tokens per document the warp sort takes

__global__ void __launch_bounds__(kDocWarps * 32) bm25_doc_pairs_kernel(
    const int* __restrict__ token_ids, const long long* __restrict__ doc_offsets, long long n_docs, int n_terms,
    long long pos_offset, int n_blocks, int* __restrict__ pair_term, int* __restrict__ pair_tf, int* __restrict__ n_unique,
    int* __restrict__ doc_len, unsigned long long* __restrict__ df, long long* __restrict__ first_pos,
    u32* __restrict__ cell, int* __restrict__ error) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  u64* keys = reinterpret_cast<u64*>(smem_raw) + static_cast<size_t>(warp) * kDocMaxLen;             // [kDocMaxLen]
  int* heads = reinterpret_cast<int*>(smem_raw + static_cast<size_t>(kDocWarps) * kDocMaxLen * 8) +
               static_cast<size_t>(warp) * (kDocMaxLen + 1);                                        // [kDocMaxLen + 1]
  for (long long doc = static_cast<long long>(blockIdx.x) * kDocWarps + warp; doc < n_docs;
       doc += static_cast<long long>(gridDim.x) * kDocWarps) {
    const long long t0 = doc_offsets[doc], t1 = doc_offsets[doc + 1];
    const int L = static_cast<int>(t1 - t0);
    if (lane == 0) doc_len[doc] = L;
    if (L > kDocMaxLen || L < 0) {
      if (lane == 0) {
        atomicExch(error, 1);
        n_unique[doc] = 0;
      }
      continue;
    }
    int p2 = 2;
    while (p2 < L) p2 <<= 1;
    // key = ~((term << 32) | position): sorting DESCENDING puts the terms in ascending order and, inside a term,
    // the positions in ascending order (the run's first entry carries the first position)
    bool bad = false;
    for (int i = lane; i < p2; i += 32) {
      u64 key = 0;  // padding sorts last
      if (i < L) {
        const int t = token_ids[t0 + i];
        bad |= (t < 0 || t >= n_terms);
        key = ~((static_cast<u64>(static_cast<u32>(t)) << 32) | static_cast<u64>(static_cast<u32>(i)));
      }
      keys[i] = key;
    }
    if (__any_sync(0xffffffffu, bad)) {
      if (lane == 0) {
        atomicExch(error, 2);
        n_unique[doc] = 0;
      }
      __syncwarp();
      continue;
    }
    __syncwarp();
    warp_sort_desc(keys, p2, lane);
    // heads of the runs of equal terms, in order
    int n_heads = 0;
    for (int base = 0; base < L; base += 32) {
      const int i = base + lane;
      bool head = false;
      if (i < L) {
        const u32 t = static_cast<u32>((~keys[i]) >> 32);
        head = (i == 0) || (static_cast<u32>((~keys[i - 1]) >> 32) != t);
      }
      const u32 m = __ballot_sync(0xffffffffu, head);
      if (head) heads[n_heads + __popc(m & ((1u << lane) - 1u))] = i;
      n_heads += __popc(m);
    }
    if (lane == 0) {
      heads[n_heads] = L;
      n_unique[doc] = n_heads;
    }
    __syncwarp();
    const int b = static_cast<int>(doc / kBuildBlk);
    for (int j = lane; j < n_heads; j += 32) {
      const int h = heads[j];
      const u64 k = ~keys[h];
      const int t = static_cast<int>(k >> 32);
      pair_term[t0 + j] = t;
      pair_tf[t0 + j] = heads[j + 1] - h;
      atomicAdd(df + t, 1ull);
      atomicMin(first_pos + t, pos_offset + t0 + static_cast<long long>(static_cast<u32>(k)));
      atomicAdd(cell + static_cast<size_t>(t) * (n_blocks + 1) + b, 1u);
    }
    __syncwarp();
  }
}

// one warp per term: cell[t][0 .. n_blocks] counts -> exclusive prefix (cell[t][b] = pairs of t in blocks < b)
__global__ void bm25_block_scan_kernel(u32* __restrict__ cell, int n_terms, int n_blocks) {
  const int lane = threadIdx.x & 31;
  const long long t = (static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
  if (t >= n_terms) return;
  u32* row = cell + static_cast<size_t>(t) * (n_blocks + 1);
  u32 carry = 0;
  for (int base = 0; base <= n_blocks; base += 32) {
    const int i = base + lane;
    const u32 v = i <= n_blocks ? row[i] : 0u;
    u32 incl = v;
#pragma unroll
    for (int m = 1; m < 32; m <<= 1) {
      const u32 o = __shfl_up_sync(0xffffffffu, incl, m);
      if (lane >= m) incl += o;
    }
    if (i <= n_blocks) row[i] = carry + incl - v;
    carry += __shfl_sync(0xffffffffu, incl, 31);
  }
}

// indptr[0 .. n] = exclusive scan of df[0 .. n) (one CTA; n is a vocabulary size)
__global__ void __launch_bounds__(1024) bm25_df_scan_kernel(const unsigned long long* __restrict__ df, int n,
                                                            long long* __restrict__ indptr) {
  __shared__ unsigned long long warp_sum[32];
  __shared__ unsigned long long carry_s;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (tid == 0) carry_s = 0;
  __syncthreads();
  for (int base = 0; base < n; base += 1024) {
    const int i = base + tid;
    const unsigned long long v = i < n ? df[i] : 0ull;
    unsigned long long incl = v;
#pragma unroll
    for (int m = 1; m < 32; m <<= 1) {
      const unsigned long long o = __shfl_up_sync(0xffffffffu, incl, m);
      if (lane >= m) incl += o;
    }
    if (lane == 31) warp_sum[warp] = incl;
    __syncthreads();
    if (warp == 0) {
      unsigned long long w = warp_sum[lane];
#pragma unroll
      for (int m = 1; m < 32; m <<= 1) {
        const unsigned long long o = __shfl_up_sync(0xffffffffu, w, m);
        if (lane >= m) w += o;
      }
      warp_sum[lane] = w;  // inclusive over warps
    }
    __syncthreads();
    const unsigned long long before = carry_s + (warp ? warp_sum[warp - 1] : 0ull);
    if (i < n) indptr[i] = static_cast<long long>(before + incl - v);
    __syncthreads();
    if (tid == 1023) carry_s = before + incl;
    __syncthreads();
  }
  if (tid == 0) indptr[n] = static_cast<long long>(carry_s);
}

// one CTA per 4096-document block, documents in order: slot = indptr[term] + offset of the block + running count
__global__ void __launch_bounds__(128) bm25_place_kernel(const long long* __restrict__ doc_offsets, long long n_docs, int n_blocks,
                                                         const int* __restrict__ pair_term, const int* __restrict__ pair_tf,
                                                         const int* __restrict__ n_unique, const long long* __restrict__ indptr,
                                                         const u32* __restrict__ cell, u32* __restrict__ cursor,
                                                         int* __restrict__ doc_ids, int* __restrict__ tf) {
  const int b = blockIdx.x, tid = threadIdx.x;
  const long long d0 = static_cast<long long>(b) * kBuildBlk, d1 = min(n_docs, d0 + kBuildBlk);
  for (long long doc = d0; doc < d1; ++doc) {
    const long long t0 = doc_offsets[doc];
    const int nu = n_unique[doc];
    for (int j = tid; j < nu; j += blockDim.x) {
      const int t = pair_term[t0 + j];
      const size_t c = static_cast<size_t>(t) * (n_blocks + 1) + b;
      const u32 r = atomicAdd(cursor + c, 1u);  // the terms of one document are distinct: no two threads share a counter
      const long long pos = indptr[t] + cell[c] + r;
      doc_ids[pos] = static_cast<int>(doc);
      tf[pos] = pair_tf[t0 + j];
    }
    __syncthreads();  // the next document's pairs must see this document's counts
    __threadfence_block();
  }
}

// ---------------------------------------------------------------------------------------------------------
__global__ void heatmap_inv_norm_kernel(const float* __restrict__ x, long long n, int d, float* __restrict__ inv) {
  const int lane = threadIdx.x & 31;
  const long long r = (static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
  if (r >= n) return;
  const float* row = x + r * d;
  float acc = 0.f;
  for (int c = lane; c < d; c += 32) acc = fmaf(row[c], row[c], acc);
#pragma unroll
  for (int m = 16; m > 0; m >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, m);
  if (lane == 0) inv[r] = sqrtf(acc) + 1e-8f;  // the divisor: X / (|X| + 1e-8)
}

__global__ void __launch_bounds__(256) heatmap_pool_kernel(const float* __restrict__ x, const float* __restrict__ div, int d,
                                                           const long long* __restrict__ bounds, double* __restrict__ pooled,
                                                           double* __restrict__ self_dot) {
  __shared__ double red[256];
  const int g = blockIdx.x, tid = threadIdx.x;
  const long long r0 = bounds[g], r1 = bounds[g + 1];
  double sq = 0.0;
  for (int c = tid; c < d; c += 256) {
    double s = 0.0;
    for (long long r = r0; r < r1; ++r) {
      const float v = x[r * d + c] / div[r];
      s += static_cast<double>(v);
      sq += static_cast<double>(v) * static_cast<double>(v);
    }
    pooled[static_cast<size_t>(g) * d + c] = s;
  }
  red[tid] = sq;
  __syncthreads();
  for (int m = 128; m > 0; m >>= 1) {
    if (tid < m) red[tid] += red[tid + m];
    __syncthreads();
  }
  if (tid == 0) self_dot[g] = red[0];
}

__global__ void __launch_bounds__(256) heatmap_gram_kernel(const double* __restrict__ pooled, const double* __restrict__ self_dot,
                                                           const long long* __restrict__ bounds, int p, int d,
                                                           float* __restrict__ tiles) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int a = blockIdx.y;
  const int b = blockIdx.x * 8 + warp;
  if (b >= p) return;
  const double* pa = pooled + static_cast<size_t>(a) * d;
  const double* pb = pooled + static_cast<size_t>(b) * d;
  double s = 0.0;
  for (int c = lane; c < d; c += 32) s = fma(pa[c], pb[c], s);
#pragma unroll
  for (int m = 16; m > 0; m >>= 1) s += __shfl_xor_sync(0xffffffffu, s, m);
  if (lane == 0) {
    const double sa = static_cast<double>(bounds[a + 1] - bounds[a]), sb = static_cast<double>(bounds[b + 1] - bounds[b]);
    if (a == b) s += sa - self_dot[a];  // the script overwrites S[i, i] (= |xn_i|^2) with 1
    double v = s / (sa * sb);
    v = v > 1.0 ? 1.0 : (v < -1.0 ? -1.0 : v);
    tiles[static_cast<size_t>(a) * p + b] = static_cast<float>(v);
  }
}

}  // namespace b2r

using namespace b2r;

extern "C" {

// Phase 1 of the postings build (everything on the device; the caller zero-initialises df and cell and fills
// first_pos with INT64_MAX): unique (term, tf) pairs per document written over the document's token slots
// (pair_term / pair_tf [total tokens]), n_unique / doc_len [n_docs], df [n_terms] (+1 per document containing the term),
// first_pos [n_terms] (min global position, pos_offset + index into token_ids), cell [n_terms][n_blocks + 1] then
// scanned in place into block offsets; indptr [n_terms + 1] = exclusive scan of df.  Documents longer than 1024 tokens or
// term ids outside [0, n_terms) fail with B2R_ERR_UNSUPPORTED / B2R_ERR_INVALID (nothing usable is left in the outputs).
int b2r_bm25_build_count_dev(const int32_t* token_ids, const int64_t* doc_offsets, int64_t n_docs, int32_t n_terms,
                             int64_t pos_offset, int32_t* pair_term, int32_t* pair_tf, int32_t* n_unique, int32_t* doc_len,
                             uint64_t* df, int64_t* first_pos, uint32_t* cell, int64_t* indptr, int32_t* error_flag,
                             void* stream) {
  if (n_docs < 1 || n_terms < 1 || !doc_offsets || !pair_term || !pair_tf || !n_unique || !doc_len || !df || !first_pos ||
      !cell || !indptr || !error_flag) {
    return fail(B2R_ERR_INVALID, "b2r_bm25_build_count_dev: bad arguments");
  }
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int n_blocks = static_cast<int>((n_docs + kBuildBlk - 1) / kBuildBlk);
  const size_t smem = static_cast<size_t>(kDocWarps) * kDocMaxLen * 8 + static_cast<size_t>(kDocWarps) * (kDocMaxLen + 1) * 4;
  B2R_CUDA(cudaFuncSetAttribute(bm25_doc_pairs_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
  const long long want = (n_docs + kDocWarps - 1) / kDocWarps;
  const int grid = static_cast<int>(std::min<long long>(want, 148LL * 16));
  bm25_doc_pairs_kernel<<<grid, kDocWarps * 32, smem, st>>>(
      token_ids, reinterpret_cast<const long long*>(doc_offsets), n_docs, n_terms, pos_offset, n_blocks, pair_term, pair_tf,
      n_unique, doc_len, reinterpret_cast<unsigned long long*>(df), reinterpret_cast<long long*>(first_pos), cell, error_flag);
  B2R_CUDA(cudaGetLastError());
  const long long scan_threads = static_cast<long long>(n_terms) * 32;
  bm25_block_scan_kernel<<<static_cast<unsigned>((scan_threads + 255) / 256), 256, 0, st>>>(cell, n_terms, n_blocks);
  B2R_CUDA(cudaGetLastError());
  bm25_df_scan_kernel<<<1, 1024, 0, st>>>(reinterpret_cast<const unsigned long long*>(df), n_terms,
                                          reinterpret_cast<long long*>(indptr));
  B2R_CUDA(cudaGetLastError());
  launch_counter() = 3;
  return B2R_OK;
}

// Phase 2: the postings.  cursor [n_terms][n_blocks + 1] zero-initialised scratch; doc_ids / tf [indptr[n_terms]].
int b2r_bm25_build_place_dev(const int64_t* doc_offsets, int64_t n_docs, int32_t n_terms, const int32_t* pair_term,
                             const int32_t* pair_tf, const int32_t* n_unique, const int64_t* indptr, const uint32_t* cell,
                             uint32_t* cursor, int32_t* doc_ids, int32_t* tf, void* stream) {
  if (n_docs < 1 || n_terms < 1 || !doc_offsets || !pair_term || !pair_tf || !n_unique || !indptr || !cell || !cursor) {
    return fail(B2R_ERR_INVALID, "b2r_bm25_build_place_dev: bad arguments");
  }
  const int n_blocks = static_cast<int>((n_docs + kBuildBlk - 1) / kBuildBlk);
  bm25_place_kernel<<<n_blocks, 128, 0, static_cast<cudaStream_t>(stream)>>>(
      reinterpret_cast<const long long*>(doc_offsets), n_docs, n_blocks, pair_term, pair_tf, n_unique,
      reinterpret_cast<const long long*>(indptr), cell, cursor, doc_ids, tf);
  B2R_CUDA(cudaGetLastError());
  launch_counter() = 1;
  return B2R_OK;
}

// Heat-map tiles: rows_dev [n, d] float32, bounds_dev [p + 1] int64 (ascending row boundaries of the p pixel groups),
// scratch: div_dev [n] float32, pooled_dev [p, d] float64, self_dot_dev [p] float64; tiles_dev [p, p] float32.
int b2r_cosine_heatmap_dev(const float* rows_dev, int64_t n, int d, const int64_t* bounds_dev, int p, float* div_dev,
                           double* pooled_dev, double* self_dot_dev, float* tiles_dev, void* stream) {
  if (n < 1 || d < 1 || p < 1 || !rows_dev || !bounds_dev || !div_dev || !pooled_dev || !self_dot_dev || !tiles_dev) {
    return fail(B2R_ERR_INVALID, "b2r_cosine_heatmap_dev: bad arguments");
  }
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  heatmap_inv_norm_kernel<<<static_cast<unsigned>((n * 32 + 255) / 256), 256, 0, st>>>(rows_dev, n, d, div_dev);
  B2R_CUDA(cudaGetLastError());
  heatmap_pool_kernel<<<p, 256, 0, st>>>(rows_dev, div_dev, d, reinterpret_cast<const long long*>(bounds_dev), pooled_dev,
                                         self_dot_dev);
  B2R_CUDA(cudaGetLastError());
  heatmap_gram_kernel<<<dim3((p + 7) / 8, p), 256, 0, st>>>(pooled_dev, self_dot_dev, reinterpret_cast<const long long*>(bounds_dev),
                                                            p, d, tiles_dev);
  B2R_CUDA(cudaGetLastError());
  launch_counter() = 3;
  return B2R_OK;
}

}  // extern "C"
