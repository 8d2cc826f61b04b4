"""Oracle for faiss.IndexFlatIP.search (faiss-cpu 1.9.0.post1), restated from its published source
(IndexFlat.cpp -> knn_inner_product -> exhaustive_inner_product_seq / _blas, utils/Heap.h).

Call sites in the reference: src/retrieval.py:622,730,1140,1165,1261,1423.

Semantics restated:
  * scores are fp32 inner products of the stored rows with each query (no normalisation);
  * per query the k best are kept with a min-heap whose top is replaced only by a STRICTLY greater
    score (CMin::cmp), so among equal scores the lower id (seen first) survives;
  * the heap is then reordered: score descending, equal scores by ascending id (cmp2);
  * k > ntotal pads the tail with id -1 and score -FLT_MAX (heap_heapify neutral element).
Summation order differs between faiss's SIMD path (nq < 20), its sgemm path (nq >= 20), numpy and a
GPU kernel, so scores agree to ~1e-6 relative, not bitwise; comparisons use assert_topk_equivalent.
"""
import numpy as np

NEG = np.float32(-3.4028235e38)


def flat_ip_scores_ref(xb, xq, dtype=np.float32):
    """All inner products, [nq, n].  dtype=np.float64 gives the exact-arithmetic reference."""
    xb = np.asarray(xb, dtype=np.float32)
    xq = np.asarray(xq, dtype=np.float32).reshape(-1, xb.shape[1])
    return xq.astype(dtype, copy=False) @ xb.astype(dtype, copy=False).T


def topk_from_scores(S, k):
    """Score descending, ties by ascending id; pads with (-FLT_MAX, -1) when k > n."""
    nq, n = S.shape
    D = np.full((nq, k), NEG, dtype=np.float32)
    I = np.full((nq, k), -1, dtype=np.int64)
    kk = min(k, n)
    if kk == 0:
        return D, I
    ids = np.arange(n)
    for qi in range(nq):
        s = S[qi]
        if kk < n:
            # every row tied with the k-th score must be a candidate before the stable sort
            kth = np.partition(s, n - kk)[n - kk]
            cand = ids[s >= kth]
        else:
            cand = ids
        order = np.lexsort((cand, -s[cand].astype(np.float64)))[:kk]
        sel = cand[order]
        D[qi, :kk] = s[sel]
        I[qi, :kk] = sel
    return D, I


def flat_ip_search_ref(xb, xq, k):
    """D, I = IndexFlatIP.search(xq, k) — fp32 scores."""
    return topk_from_scores(flat_ip_scores_ref(xb, xq, np.float32), int(k))


def assert_topk_equivalent(D, I, S_ref, k, rtol=1e-5, atol=1e-6, what="top-k", pad_score=NEG):
    """Check a (D, I) result against full reference scores S_ref [nq, n] (fp32 or fp64).

    north_star's bar: identical top-k ids except ties within tolerance, scores within 1e-5 relative.
      1. every returned id is valid and unique, and its returned score matches S_ref[id] (rtol);
      2. rows are sorted by score descending;
      3. the returned set equals the reference top-k set except for ids whose reference score lies
         within the tolerance band around the reference k-th score (tie / rounding swaps).
    """
    D = np.asarray(D)
    I = np.asarray(I)
    S_ref = np.asarray(S_ref, dtype=np.float64)
    nq, n = S_ref.shape
    assert D.shape == (nq, k) and I.shape == (nq, k), f"{what}: bad result shape {D.shape} {I.shape}"
    kk = min(k, n)
    for qi in range(nq):
        ids = I[qi, :kk]
        assert np.all((ids >= 0) & (ids < n)), f"{what}: q{qi} invalid id in {ids}"
        assert len(set(ids.tolist())) == kk, f"{what}: q{qi} duplicate ids"
        assert np.all(I[qi, kk:] == -1), f"{what}: q{qi} padding ids not -1"
        assert np.all(D[qi, kk:] == pad_score), f"{what}: q{qi} padding scores are not the pad value"
        ref_scores = S_ref[qi, ids]
        tol = atol + rtol * np.abs(ref_scores)
        assert np.all(np.abs(D[qi, :kk] - ref_scores) <= tol), (
            f"{what}: q{qi} score mismatch, max err {np.max(np.abs(D[qi, :kk] - ref_scores))}")
        assert np.all(np.diff(D[qi, :kk]) <= 0), f"{what}: q{qi} scores not sorted descending"
        if kk == n:
            continue
        part = np.partition(S_ref[qi], n - kk)
        kth = part[n - kk]
        band = atol + rtol * abs(kth)
        must_have = np.nonzero(S_ref[qi] > kth + band)[0]
        missing = set(must_have.tolist()) - set(ids.tolist())
        assert not missing, f"{what}: q{qi} missing ids {sorted(missing)[:5]} that beat the k-th score"
        assert np.all(ref_scores >= kth - band), f"{what}: q{qi} returned an id below the k-th score band"
