/*
 * b2r.h — C ABI of the B200-native retrieval hot path (libb2r.so).
 *
 * This is the drop-in boundary for the retrieval arithmetic of HEIcby/RAG-Challenge-2.
 * The reference has no FFI of its own: its hot path is two third-party Python objects,
 * faiss.IndexFlatIP (faiss-cpu==1.9.0.post1) and rank_bm25.BM25Okapi (rank-bm25==0.2.2),
 * called from src/retrieval.py.  Each entry point below names the reference call site it
 * replaces (paths relative to the reference checkout).  The Python classes in
 * rag-challenge-2_b200/ bind these symbols with ctypes and mirror the faiss / rank_bm25 /
 * retriever interfaces on top of them.
 *
 * Conventions
 *   - every function returns 0 on success or a negative B2R_ERR_* code; no exceptions cross
 *     the ABI; b2r_last_error() returns a thread-local message for the last failure.
 *   - plain pointers and sizes only.  "dev" pointers are CUDA device pointers on the handle's
 *     device, "host" pointers are ordinary host memory.  `stream` is a cudaStream_t passed as
 *     void* (NULL = the legacy default stream).
 *   - handles own their device memory unless created by an *_adopt call.
 *   - *_host entry points are synchronous and may be called concurrently from many host
 *     threads on one handle (src/retrieval.py:1498-1504 does exactly that with faiss);
 *     *_dev entry points only enqueue work on `stream`.
 *   - ids are int64 row numbers in insertion order (faiss semantics); a missing result is
 *     id -1 with score -3.4028235e38 (faiss pads k > ntotal the same way).
 *   - result order: score descending, ties by ascending id.
 */
#ifndef B2R_H_
#define B2R_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define B2R_OK 0
#define B2R_ERR_INVALID (-1)     /* bad argument (NULL handle, d mismatch, k out of range ...) */
#define B2R_ERR_CUDA (-2)        /* a CUDA runtime call failed; message has the CUDA error    */
#define B2R_ERR_NOMEM (-3)       /* device or host allocation failed                          */
#define B2R_ERR_UNSUPPORTED (-4) /* valid request outside what the kernels implement          */

#define B2R_MAX_K 1024 /* largest k of the fused select (faiss-gpu caps at 2048)          */
#define B2R_LAUNCH_OVERLAP_PREVIOUS 1u /* launch flag of the *_ex entry points, see b2r_flat_search_dev_ex */

typedef struct b2r_flat b2r_flat; /* exact inner-product index  == faiss.IndexFlatIP          */
typedef struct b2r_bm25 b2r_bm25; /* BM25Okapi postings index   == rank_bm25.BM25Okapi        */

/* ---- library ---------------------------------------------------------------------------- */
const char* b2r_last_error(void);
int b2r_version(void);
int b2r_device_count(int* out);
/* Tunables for experiments and tests (pipeline depth, select buffer size ...).  Unknown keys
 * return B2R_ERR_INVALID.  Not part of the reference surface. */
int b2r_tunable_set(const char* key, int value);
int b2r_tunable_get(const char* key, int* value);

/* ---- flat inner-product index ------------------------------------------------------------ */
/* faiss.IndexFlatIP(d)                                   src/ingestion.py:82 */
int b2r_flat_create(int d, int device, b2r_flat** out);
/* index.add(float32[n,d]) — copies the rows              src/ingestion.py:83
 * rows_on_device != 0: `rows` is a device pointer on the handle's device. */
int b2r_flat_add(b2r_flat* h, const float* rows, int64_t n, int rows_on_device);
/* Zero-copy variant for corpora generated or loaded straight into HBM: the handle borrows
 * dev_rows[n,d] (row-major, 16-byte aligned); the caller keeps it alive.  Index must be empty. */
int b2r_flat_adopt(b2r_flat* h, const float* dev_rows, int64_t n);
/* index.ntotal / index.d                                 generate_similarity_matrix.py:49-51 */
int b2r_flat_ntotal(const b2r_flat* h, int64_t* out);
int b2r_flat_dim(const b2r_flat* h, int* out);
/* Device pointer of the row storage (for the N x N cosine analysis and peer access). */
int b2r_flat_rows_dev(const b2r_flat* h, const float** out);
/* D, I = index.search(x=float32[nq,d], k=k)              src/retrieval.py:622,730,1140,1165,1261,1423
 * q, D[nq*k], I[nq*k] are host buffers.  Synchronous, thread-safe. */
int b2r_flat_search_host(b2r_flat* h, const float* q, int nq, int k, float* D, int64_t* I);
/* Concurrent callers of b2r_flat_search_host on one index are combined: requests that arrive while a search is in
 * flight are run together as one batch (one pass over the rows for all of them).  Counts batches / requests so far. */
int b2r_flat_combine_stats(b2r_flat* h, int64_t* batches, int64_t* requests);
/* Same search with everything resident: q_dev[nq,d], D_dev[nq,k], I_dev[nq,k] on the device,
 * enqueued on `stream`.  id_offset is added to every valid id (global id of a row shard). */
int b2r_flat_search_dev(b2r_flat* h, const float* q_dev, int nq, int k, float* D_dev,
                        int64_t* I_dev, int64_t id_offset, void* stream);
/* Same, with launch flags.  B2R_LAUNCH_OVERLAP_PREVIOUS (defined below): a search of one or two queries (the scan
 * kernel) is launched with programmatic stream serialization — its pass over the rows may start while the previous
 * kernel of `stream` is still finishing, once that kernel has issued griddepcontrol.launch_dependents (this library's
 * fusion / exchange kernels do so at their start; any other kernel releases it by exiting, i.e. ordinary stream
 * order).  The scan writes nothing before the previous kernel has completed (griddepcontrol.wait ahead of its
 * epilogue), so D_dev / I_dev and the workspace may be the ones that kernel still reads; q_dev must have been
 * complete when the previous kernel STARTED.  A stream of hybrid steps uses this to hide the fusion / exchange kernel
 * of step s — including its wait for the slowest rank — behind the first tiles of step s + 1. */
int b2r_flat_search_dev_ex(b2r_flat* h, const float* q_dev, int nq, int k, float* D_dev,
                           int64_t* I_dev, int64_t id_offset, void* stream, unsigned flags);
/* index.reconstruct(i) / index.reconstruct_batch(ids)    src/retrieval.py:573;
 * generate_similarity_matrix.py:61.  out is host float32[m,d]; any id outside [0,ntotal)
 * fails with B2R_ERR_INVALID (faiss raises; the reference catches it, :574-576). */
int b2r_flat_reconstruct(b2r_flat* h, const int64_t* ids, int m, float* out);
/* Optional: keep a bfloat16 copy of the rows (+50 % memory) for the tensor-core candidate pass of batches of more
 * than 128 queries — half the bytes per row and twice the MMA rate.  Results stay exact: the candidates (128 per query)
 * are rescored in fp32 and the answer is only accepted when the bf16 error bound proves the candidate set complete;
 * otherwise the query is redone with 3xTF32.  The copy is (re)built lazily by the next batched search. */
int b2r_flat_set_candidate_bf16(b2r_flat* h, int enable);
int b2r_flat_destroy(b2r_flat* h);
/* Number of kernel launches the last search on this thread enqueued (bench accounting). */
int b2r_last_launch_count(void);

/* ---- top-k merge --------------------------------------------------------------------------- */
/* Merge n_lists per-shard / per-database result lists into one top-k per query; replaces the
 * Python sorted(...)[:top_n] over per-report hits (src/retrieval.py:293-295,1719,1754) and is
 * the consumer of the multi-GPU all-gather.  Layout: D_lists[n_lists][nq][k_in] (list-major,
 * i.e. exactly what an all-gather of per-rank [nq,k_in] results produces), same for I_lists.
 * Entries with id < 0 are ignored.  Output D[nq,k_out], I[nq,k_out], padded with -1. */
int b2r_topk_merge_dev(const float* D_lists, const int64_t* I_lists, int n_lists, int nq, int k_in,
                       int k_out, float* D, int64_t* I, void* stream);
/* float64-score variant (BM25 lists). */
int b2r_topk_merge_f64_dev(const double* S_lists, const int64_t* I_lists, int n_lists, int nq,
                           int k_in, int k_out, double* S, int64_t* I, void* stream);
int b2r_topk_merge_host(const float* D_lists, const int64_t* I_lists, int n_lists, int nq,
                        int k_in, int k_out, float* D, int64_t* I, int device);

/* ---- BM25Okapi ----------------------------------------------------------------------------- */
/* BM25Okapi(corpus) after host tokenisation              src/ingestion.py:19-22
 * The corpus arrives as term-major CSR postings (host pointers): for term t the postings are
 * doc_ids[indptr[t]..indptr[t+1]) ascending, with term frequencies tf[...].  doc_len[n_docs] are
 * the token counts.  idf[n_terms] (float64) and avgdl are the *corpus-global* statistics
 * computed by the caller exactly as rank_bm25 does (epsilon floor included), so that a document
 * shard scores identically to the unsharded corpus.  The library precomputes the per-posting
 * weight tf*(k1+1)/(tf+k1*(1-b+b*dl/avgdl)) in float64 and keeps it in float64 (12 B / posting):
 * scores are bit-identical to rank_bm25's float64 arithmetic. */
int b2r_bm25_create(int64_t n_docs, int32_t n_terms, const int64_t* indptr, const int32_t* doc_ids,
                    const int32_t* tf, const int32_t* doc_len, double avgdl, const double* idf,
                    double k1, double b, int device, b2r_bm25** out);
/* Extended form: inputs_on_device != 0 takes all array arguments as device pointers (corpora built
 * in HBM; the arrays are only read during the call); weight_bits = 64 keeps float64 weights
 * (12 B / posting, bit-identical to rank_bm25; what b2r_bm25_create does), 32 is the explicit opt-in
 * that rounds the stored weights to float32 (8 B / posting; scores within 1e-6 relative). */
int b2r_bm25_create_ex(int64_t n_docs, int32_t n_terms, const int64_t* indptr, const int32_t* doc_ids,
                       const int32_t* tf, const int32_t* doc_len, double avgdl, const double* idf,
                       double k1, double b, int device, int inputs_on_device, int weight_bits,
                       b2r_bm25** out);
int b2r_bm25_ndocs(const b2r_bm25* h, int64_t* out);
/* number of postings and the stored weight width (bench accounting: bytes per posting) */
int b2r_bm25_postings(const b2r_bm25* h, int64_t* n_post, int* weight_bits);
/* scores = bm25.get_scores(tokens)                       src/retrieval.py:262
 * One query: term_ids[n_q_terms] (duplicates kept, unknown tokens dropped by the caller);
 * scores is host float64[n_docs].  Parity/debug path: writes the dense vector. */
int b2r_bm25_scores_host(b2r_bm25* h, const int32_t* term_ids, int n_q_terms, double* scores);
/* get_scores + sorted(range(n), key=scores.__getitem__, reverse=True)[:k]
 *                                                        src/retrieval.py:262-265
 * Batched and fused: query i owns term_ids[term_offsets[i]..term_offsets[i+1]).  No dense score
 * vector is written.  S[nq,k] float64 scores, I[nq,k] doc ids; ties (including the zero-score
 * tail) keep ascending doc id, as Python's stable sort does. */
int b2r_bm25_topk_host(b2r_bm25* h, const int32_t* term_ids, const int32_t* term_offsets, int nq,
                       int k, double* S, int64_t* I);
int b2r_bm25_topk_dev(b2r_bm25* h, const int32_t* term_ids_dev, const int32_t* term_offsets_dev,
                      const int32_t* term_offsets_host, int nq, int k, double* S_dev,
                      int64_t* I_dev, int64_t id_offset, void* stream);
/* Same, with launch flags.  B2R_LAUNCH_OVERLAP_PREVIOUS: the kernel is launched with programmatic stream
 * serialization, so it may run BESIDE the previous kernel of `stream` once all CTAs of that kernel are
 * running (the flat scan kernels issue griddepcontrol.launch_dependents at their start; any other kernel
 * releases it when it exits, i.e. ordinary stream order).  Only for work that does not read what the previous
 * kernel writes — the hybrid path launches BM25 this way right behind the dense scan. */
#ifndef B2R_LAUNCH_OVERLAP_PREVIOUS
#define B2R_LAUNCH_OVERLAP_PREVIOUS 1u
#endif
int b2r_bm25_topk_dev_ex(b2r_bm25* h, const int32_t* term_ids_dev, const int32_t* term_offsets_dev,
                         const int32_t* term_offsets_host, int nq, int k, double* S_dev,
                         int64_t* I_dev, int64_t id_offset, void* stream, unsigned flags);
int b2r_bm25_destroy(b2r_bm25* h);

/* ---- dense + BM25 fusion ------------------------------------------------------------------- */
/* Hybrid top-k for BASELINE config 4.  The reference defines no dense+BM25 formula (PARITY
 * UNPINNED); this applies its aggregation rule src/retrieval.py:1639-1661 with BM25
 * as one more retrieval source: base = dense score if the chunk is a dense hit, else the BM25 score
 * min-max normalised over the query's BM25 hits; final = base * 1.1 when both retrievers found it.
 * D/I[nq,k_dense] dense hits, S/J[nq,k_bm25] BM25 hits (ids must be < 2^32), outputs F/O[nq,k_out]
 * (score descending, id ascending) and optional methods[nq,k_out] (bit 0 dense, bit 1 BM25). */
int b2r_hybrid_fuse_dev(const float* D, const int64_t* I, const double* S, const int64_t* J, int nq,
                        int k_dense, int k_bm25, int k_out, float* F, int64_t* O,
                        unsigned char* methods, void* stream);

/* Multi-GPU variant: exchange + merge + fusion in ONE kernel over NVLink peer memory (replaces the
 * all-gather, two merges and the fusion of the NCCL path).  Every rank passes its local hits
 * D/I/S/J[nq,k] (global ids) and the device addresses of all ranks' symmetric buffers
 * (peer_bufs[world], each b2r_hybrid_exchange_bytes(nq_max, k_max, world) bytes, zero-initialised,
 * mapped for peer access); `step` (>= 1) must increase by one per call, in lockstep on all ranks.
 * The kernel stores the hits into every rank's buffer, publishes system-scope flags, waits for the
 * peers' flags and merges + fuses; all ranks obtain the same F/O[nq,k].
 * The wait is bounded (tunable "exchange.timeout_ms", default 10 s): if a peer never publishes this step
 * (it failed before launching, or counts steps differently) the kernel stores 1 + that peer's rank into
 * *status (a zero-initialised device-visible word the caller owns and checks after synchronising; the
 * query's outputs are then padding) — or traps, surfacing as a CUDA error, when status is NULL. */
size_t b2r_hybrid_exchange_bytes(int nq_max, int k_max, int world);
int b2r_hybrid_exchange_fuse_dev(const float* D, const int64_t* I, const double* S, const int64_t* J,
                                 int nq, int k, void* const* peer_bufs, int rank, int world, int nq_max,
                                 int k_max, uint32_t step, float* F, int64_t* O, uint32_t* status, void* stream);

/* ---- one hybrid step behind one call (host buffers) ------------------------------------------- */
/* What the drop-in does per question on one index: q[nq,d] + token ids in, fused top-k out.  Replaces the
 * Python sequence index.search -> bm25.get_scores -> sorted(...)[:top_n] (src/retrieval.py:262-295,1140-1157)
 * plus the fusion above.  The context owns a stream and pinned / device staging (grown on demand); steps on one
 * context are serialised — create one context per host thread.  Inside: ONE pinned host->device copy of the
 * packed inputs, the dense scan, BM25 launched beside it (B2R_LAUNCH_OVERLAP_PREVIOUS), the fusion kernel — or,
 * when world > 1, the one-kernel exchange + merge + fusion over peer_bufs (see above; `step` as there) — the
 * device->host copies of F / O and one stream synchronisation.  id_offset: global id of local row / doc 0. */
typedef struct b2r_hybrid_ctx b2r_hybrid_ctx;
int b2r_hybrid_ctx_create(int device, b2r_hybrid_ctx** out);
int b2r_hybrid_ctx_destroy(b2r_hybrid_ctx* ctx);
/* bytes moved host->device and device->host by the last step on this context (bench accounting) */
int b2r_hybrid_ctx_bytes(const b2r_hybrid_ctx* ctx, size_t* h2d, size_t* d2h);
int b2r_hybrid_search_host(b2r_hybrid_ctx* ctx, b2r_flat* flat, b2r_bm25* bm25, const float* q, int nq, int d,
                           const int32_t* term_ids, const int32_t* term_offsets, int k, int64_t id_offset,
                           void* const* peer_bufs, int rank, int world, int nq_max, int k_max, uint32_t step,
                           float* F, int64_t* O);
/* The same call in two halves, so that a caller with several questions pending (the reference answers questions from a
 * thread pool, src/questions_processing.py:1123-1125) keeps the GPU busy: submit copies the inputs into one of the
 * context's two staging slots, enqueues the H2D copy and the step and returns a ticket; collect waits for that step and
 * copies F / O out.  At most two submissions may be outstanding per context; they execute in submission order on the
 * context's stream, and with world > 1 every rank must submit the same sequence.  search_host == submit + collect. */
int b2r_hybrid_submit_host(b2r_hybrid_ctx* ctx, b2r_flat* flat, b2r_bm25* bm25, const float* q, int nq, int d,
                           const int32_t* term_ids, const int32_t* term_offsets, int k, int64_t id_offset,
                           void* const* peer_bufs, int rank, int world, int nq_max, int k_max, uint32_t step,
                           int* ticket);
int b2r_hybrid_collect_host(b2r_hybrid_ctx* ctx, int ticket, float* F, int64_t* O);

/* The same step with everything resident (device pointers; enqueued on `stream`, nothing synchronised): the dense scan
 * and BM25 launched beside it; for up to 4 queries and k <= 256 the LAST CTA of each query's BM25 grid waits for the
 * scan (griddepcontrol.wait) and fuses — or, world > 1, exchanges + merges + fuses over peer_bufs — in place: two
 * launches per step, no kernel launch between the end of the scan and the result.  Otherwise: scan, BM25, fusion /
 * exchange kernel.  D/I [nq,k] dense hits, S/J [nq,k] BM25 hits, F/O [nq,k] fused results.  status: see
 * b2r_hybrid_exchange_fuse_dev (may be NULL).  flags: B2R_LAUNCH_OVERLAP_PREVIOUS chains the step behind the previous
 * step on this stream — its scan queues up behind the previous scan and streams while the previous step's tail
 * (fusion, exchange, the wait for the slowest rank) completes; q_dev and the term buffers must not be produced by the
 * previous step.  launches (optional): kernels enqueued. */
int b2r_hybrid_step_dev(b2r_flat* flat, b2r_bm25* bm25, const float* q_dev, int nq,
                        const int32_t* term_ids_dev, const int32_t* term_offsets_dev, int k,
                        int64_t id_offset, void* const* peer_bufs, int rank, int world, int nq_max,
                        int k_max, uint32_t step, uint32_t* status, float* D_dev, int64_t* I_dev,
                        double* S_dev, int64_t* J_dev, float* F_dev, int64_t* O_dev, void* stream,
                        unsigned flags, int* launches);

/* ---- N x N cosine analysis ---------------------------------------------------------------- */
/* compute_cosine_similarity_matrix + off-diagonal statistics
 *                             semantic_similarity_analysis/generate_similarity_matrix.py:84-98,162-171
 * rows_dev[n,d] float32 (d a multiple of 32).  Never materialises the n x n matrix: streams tiles of
 * S = Xn Xn^T (Xn = X / (|X| + 1e-8), clipped to [-1,1]) through the tensor cores (3xTF32) and
 * reduces them.  Uses the symmetry of S: the pairs (i, j) with i < j and j in [row_begin, row_end)
 * are visited once and counted twice, so ranges covering [0, n) add up to all off-diagonal entries
 * (the unit of sharding across GPUs; work grows with j).
 * stats[0..5] = off-diagonal sum, sum of squares, min, max, count, (reserved);
 * hist[n_bins] (n_bins <= 8192) counts the off-diagonal entries over [-1,1] (for the median).
 * Synchronises `stream` before returning (scratch is freed).  k must be 0: per-row nearest
 * neighbours are b2r_flat_search_dev on the normalised rows. */
int b2r_cosine_stats_dev(const float* rows_dev, int64_t n, int d, int64_t row_begin, int64_t row_end,
                         double* stats_dev, unsigned long long* hist_dev, int n_bins, int k,
                         float* topk_S_dev, int64_t* topk_I_dev, int device, void* stream);

/* ---- index build on the GPU ------------------------------------------------------------------ */
/* BM25Okapi([chunk.split() ...]) after host tokenisation, from a token stream       src/ingestion.py:19-22
 * token_ids[total] (term ids of all documents back to back), doc_offsets[n_docs + 1]; all pointers are device
 * pointers, everything is enqueued on `stream`.  Phase 1 writes every document's unique (term, tf) pairs over its own
 * token slots (pair_term / pair_tf [total]), n_unique / doc_len [n_docs], df [n_terms] (caller zeroes it),
 * first_pos [n_terms] (caller fills it with INT64_MAX; receives pos_offset + the index of every term's first token —
 * rank_bm25 averages the idf values in first-seen order), cell [n_terms][ceil(n_docs / 4096) + 1] (caller zeroes it;
 * left holding every term's posting offset of every 4096-document block), indptr [n_terms + 1], and *error_flag
 * (1: a document has more than 1024 tokens, 2: a term id is out of range — the outputs are then unusable).
 * Phase 2 places the postings: doc_ids / tf [indptr[n_terms]], ascending document ids inside a term; cursor is a
 * zeroed scratch shaped like cell. */
int b2r_bm25_build_count_dev(const int32_t* token_ids, const int64_t* doc_offsets, int64_t n_docs,
                             int32_t n_terms, int64_t pos_offset, int32_t* pair_term, int32_t* pair_tf,
                             int32_t* n_unique, int32_t* doc_len, uint64_t* df, int64_t* first_pos,
                             uint32_t* cell, int64_t* indptr, int32_t* error_flag, void* stream);
int b2r_bm25_build_place_dev(const int64_t* doc_offsets, int64_t n_docs, int32_t n_terms,
                             const int32_t* pair_term, const int32_t* pair_tf, const int32_t* n_unique,
                             const int64_t* indptr, const uint32_t* cell, uint32_t* cursor,
                             int32_t* doc_ids, int32_t* tf, void* stream);
/* Down-sampled heat map of the N x N cosine matrix      generate_similarity_matrix.py:124-160 (plots the full matrix)
 * bounds[p + 1]: ascending row boundaries of the p pixel groups; tiles[p, p] = mean of clip(Xn Xn^T) over the rows of
 * group a and the columns of group b, the diagonal of the matrix taken as 1.  Scratch: div[n] float32, pooled[p, d] and
 * self_dot[p] float64. */
int b2r_cosine_heatmap_dev(const float* rows_dev, int64_t n, int d, const int64_t* bounds_dev, int p,
                           float* div_dev, double* pooled_dev, double* self_dot_dev, float* tiles_dev,
                           void* stream);

#ifdef __cplusplus
}
#endif
#endif /* B2R_H_ */
