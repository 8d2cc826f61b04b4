/* A host program in plain C (C99) over include/b2r.h — what a non-Python caller of the retrieval path binds.
 * Built and run by tests/test_abi_c.py:  gcc -std=c99 -Wall -Werror -Iinclude consumer.c -L<pkg> -lb2r
 *   no GPU : every create call must fail with an error message (there is no CPU fallback)   -> exit 0
 *   GPU    : IndexFlatIP(d).add / search / reconstruct (src/ingestion.py:82-83, src/retrieval.py:1140,573),
 *            BM25Okapi postings + fused top-k (src/retrieval.py:262-265), list merge, dense + BM25 fusion and the
 *            one-call hybrid step, each checked against values computed here in C                 -> exit 0
 * Prints one line per check; a failed check exits 1. */
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "b2r.h"

#define CHECK(cond, what)                                                             \
  do {                                                                                \
    if (!(cond)) {                                                                    \
      fprintf(stderr, "FAILED: %s (%s) [%s]\n", what, #cond, b2r_last_error());       \
      return 1;                                                                       \
    }                                                                                 \
    printf("ok: %s\n", what);                                                         \
  } while (0)

static uint32_t rng_state = 12345u;
static float frand(void) { /* xorshift, uniform in [-1, 1) */
  rng_state ^= rng_state << 13;
  rng_state ^= rng_state >> 17;
  rng_state ^= rng_state << 5;
  return (float)(rng_state >> 8) / 8388608.0f - 1.0f;
}

int main(void) {
  int n_dev = -1;
  CHECK(b2r_version() == 100, "b2r_version");
  int rc = b2r_device_count(&n_dev);
  b2r_flat* flat = NULL;
  if (rc != B2R_OK || n_dev <= 0) {
    CHECK(b2r_flat_create(64, 0, &flat) < 0 && flat == NULL, "no GPU: b2r_flat_create fails");
    CHECK(strlen(b2r_last_error()) > 0, "no GPU: error message set");
    printf("done (no GPU)\n");
    return 0;
  }

  /* ---- flat index: n rows of dimension d, query = row 17 scaled by 2 ---------------------------------------- */
  enum { N = 3000, D = 64, K = 5, NQ = 2 };
  float* rows = (float*)malloc(sizeof(float) * N * D);
  for (int i = 0; i < N * D; ++i) rows[i] = frand();
  CHECK(b2r_flat_create(D, 0, &flat) == B2R_OK, "b2r_flat_create");
  CHECK(b2r_flat_add(flat, rows, 1000, 0) == B2R_OK && b2r_flat_add(flat, rows + 1000 * D, N - 1000, 0) == B2R_OK,
        "b2r_flat_add (two batches)");
  int64_t ntotal = 0;
  int dim = 0;
  CHECK(b2r_flat_ntotal(flat, &ntotal) == B2R_OK && ntotal == N && b2r_flat_dim(flat, &dim) == B2R_OK && dim == D,
        "ntotal / d");
  float q[NQ * D];
  for (int j = 0; j < D; ++j) {
    q[j] = 2.0f * rows[17 * D + j];
    q[D + j] = frand();
  }
  float Dres[NQ * K];
  int64_t Ires[NQ * K];
  CHECK(b2r_flat_search_host(flat, q, NQ, K, Dres, Ires) == B2R_OK, "b2r_flat_search_host");
  /* brute force in C (double accumulation): best K for both queries */
  for (int qi = 0; qi < NQ; ++qi) {
    double best[K];
    int64_t bid[K];
    for (int s = 0; s < K; ++s) { best[s] = -1e300; bid[s] = -1; }
    for (int i = 0; i < N; ++i) {
      double acc = 0.0;
      for (int j = 0; j < D; ++j) acc += (double)rows[i * D + j] * (double)q[qi * D + j];
      int s = K;
      while (s > 0 && acc > best[s - 1]) --s;
      if (s < K) {
        for (int t = K - 1; t > s; --t) { best[t] = best[t - 1]; bid[t] = bid[t - 1]; }
        best[s] = acc;
        bid[s] = i;
      }
    }
    for (int s = 0; s < K; ++s) {
      if (Ires[qi * K + s] != bid[s] || fabs(Dres[qi * K + s] - best[s]) > 1e-5 * fabs(best[s]) + 1e-6) {
        fprintf(stderr, "FAILED: search q%d slot %d: got (%lld, %g) want (%lld, %g)\n", qi, s,
                (long long)Ires[qi * K + s], Dres[qi * K + s], (long long)bid[s], best[s]);
        return 1;
      }
    }
  }
  printf("ok: search == brute force in C (ids, scores 1e-5)\n");
  CHECK(Ires[0] == 17, "row 17 finds itself");
  int64_t ids[2] = {17, N - 1};
  float back[2 * D];
  CHECK(b2r_flat_reconstruct(flat, ids, 2, back) == B2R_OK && memcmp(back, rows + 17 * D, sizeof(float) * D) == 0 &&
            memcmp(back + D, rows + (N - 1) * D, sizeof(float) * D) == 0,
        "b2r_flat_reconstruct returns the stored bytes");
  int64_t bad = N;
  CHECK(b2r_flat_reconstruct(flat, &bad, 1, back) == B2R_ERR_INVALID, "reconstruct out of range -> B2R_ERR_INVALID");
  float big_D[8];
  int64_t big_I[8];
  b2r_flat* tiny = NULL;
  CHECK(b2r_flat_create(D, 0, &tiny) == B2R_OK && b2r_flat_add(tiny, rows, 3, 0) == B2R_OK &&
            b2r_flat_search_host(tiny, q, 1, 8, big_D, big_I) == B2R_OK && big_I[2] >= 0 && big_I[3] == -1 &&
            big_I[7] == -1 && big_D[7] < -3.0e38f,
        "k > ntotal pads with (-FLT_MAX, -1)");
  b2r_flat_destroy(tiny);

  /* ---- BM25: 4 documents, 3 terms, postings by hand ------------------------------------------------------------
   * doc0 = a a b, doc1 = b c, doc2 = a, doc3 = c c c c          term ids a=0 b=1 c=2 */
  const int64_t indptr[4] = {0, 2, 4, 6};
  const int32_t doc_ids[6] = {0, 2, 0, 1, 1, 3};
  const int32_t tf[6] = {2, 1, 1, 1, 1, 4};
  const int32_t doc_len[4] = {3, 2, 1, 4};
  const double avgdl = 10.0 / 4.0, k1 = 1.5, b = 0.75;
  double idf[3];
  for (int t = 0; t < 3; ++t) idf[t] = log(4 - 2 + 0.5) - log(2 + 0.5); /* df = 2 for every term: idf = 0 */
  idf[0] = 0.7;                                                          /* caller-supplied statistics: any values */
  idf[1] = 0.2;
  idf[2] = 1.1;
  b2r_bm25* bm = NULL;
  CHECK(b2r_bm25_create_ex(4, 3, indptr, doc_ids, tf, doc_len, avgdl, idf, k1, b, 0, 0, 64, &bm) == B2R_OK,
        "b2r_bm25_create_ex (float64 weights)");
  const int32_t terms[3] = {0, 2, 0}; /* query "a c a": the duplicated token counts twice */
  double scores[4], want[4] = {0, 0, 0, 0};
  CHECK(b2r_bm25_scores_host(bm, terms, 3, scores) == B2R_OK, "b2r_bm25_scores_host");
  for (int i = 0; i < 3; ++i) {
    int t = terms[i];
    for (int64_t p = indptr[t]; p < indptr[t + 1]; ++p) {
      double f = tf[p], dl = doc_len[doc_ids[p]];
      want[doc_ids[p]] += idf[t] * (f * (k1 + 1) / (f + k1 * (1 - b + b * dl / avgdl)));
    }
  }
  CHECK(memcmp(scores, want, sizeof(want)) == 0, "BM25 scores bit-identical to the formula evaluated in C");
  const int32_t offs[2] = {0, 3};
  double S[3];
  int64_t J[3];
  CHECK(b2r_bm25_topk_host(bm, terms, offs, 1, 3, S, J) == B2R_OK, "b2r_bm25_topk_host");
  int order_ok = 1;
  for (int s = 0; s < 3; ++s) {
    order_ok &= (S[s] == want[J[s]]);
    if (s) order_ok &= (S[s - 1] > S[s] || (S[s - 1] == S[s] && J[s - 1] < J[s]));
  }
  CHECK(order_ok, "BM25 top-3 sorted (score desc, id asc) with exact scores");

  /* ---- merge of two shard lists, fusion, one-call hybrid step -------------------------------------------------- */
  float Dl[2 * 1 * 2] = {0.9f, 0.5f, 0.9f, 0.7f}; /* [n_lists=2][nq=1][k_in=2] */
  int64_t Il[2 * 1 * 2] = {4, 9, 1, 3};
  float Dm[3];
  int64_t Im[3];
  CHECK(b2r_topk_merge_host(Dl, Il, 2, 1, 2, 3, Dm, Im, 0) == B2R_OK && Im[0] == 4 && Im[1] == 1 && Im[2] == 3 &&
            Dm[2] == 0.7f,
        "b2r_topk_merge_host (equal scores: list order = ascending global id of row shards)");
  b2r_hybrid_ctx* ctx = NULL;
  CHECK(b2r_hybrid_ctx_create(0, &ctx) == B2R_OK, "b2r_hybrid_ctx_create");
  /* a 4-row index aligned with the 4 BM25 documents */
  b2r_flat* four = NULL;
  CHECK(b2r_flat_create(D, 0, &four) == B2R_OK && b2r_flat_add(four, rows, 4, 0) == B2R_OK, "4-row index");
  float F[4];
  int64_t O[4];
  CHECK(b2r_hybrid_search_host(ctx, four, bm, q + D, 1, D, terms, offs, 4, 0, NULL, 0, 1, 0, 0, 0, F, O) == B2R_OK,
        "b2r_hybrid_search_host");
  int seen = 0;
  for (int s = 0; s < 4; ++s) {
    if (O[s] >= 0 && O[s] < 4) seen |= 1 << O[s];
    if (s && O[s] >= 0) { if (!(F[s - 1] >= F[s])) seen = -1; }
  }
  CHECK(seen == 15, "hybrid step returns the 4 chunks once each, scores descending");
  size_t h2d = 0, d2h = 0;
  CHECK(b2r_hybrid_ctx_bytes(ctx, &h2d, &d2h) == B2R_OK && h2d >= sizeof(float) * D && d2h == 12u * 4u,
        "b2r_hybrid_ctx_bytes");
  b2r_hybrid_ctx_destroy(ctx);
  b2r_flat_destroy(four);
  b2r_bm25_destroy(bm);
  b2r_flat_destroy(flat);
  free(rows);
  printf("done (GPU)\n");
  return 0;
}
