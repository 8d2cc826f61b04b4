"""CPU restatement of (a) the top-k list merge and (b) the dense + BM25 fusion.  TEST INFRASTRUCTURE.

(a) restates what the reference does in Python when it merges per-report hits:
    all_results.sort(key=score, reverse=True)[:top_n] (src/retrieval.py:293-295, 1719, 1754) —
    a stable descending sort, so equal scores keep list order.
(b) PARITY UNPINNED: the reference has no dense+BM25 formula.  It applies the reference's
    aggregation rule (src/retrieval.py:1639-1661) with BM25 as one more retrieval source; see
    csrc/hybrid_fuse.cu for the definition this function restates in numpy float64.
"""
import numpy as np


def merge_lists_ref(D_lists, I_lists, k_out, pad_score):
    """D_lists/I_lists: [n_lists, nq, k_in]; stable descending merge in (list, slot) order."""
    D_lists = np.asarray(D_lists)
    I_lists = np.asarray(I_lists)
    n_lists, nq, k_in = D_lists.shape
    D = np.full((nq, k_out), pad_score, dtype=D_lists.dtype)
    I = np.full((nq, k_out), -1, dtype=np.int64)
    for q in range(nq):
        cand = [(D_lists[l, q, s], l * k_in + s, I_lists[l, q, s])
                for l in range(n_lists) for s in range(k_in) if I_lists[l, q, s] >= 0]
        cand.sort(key=lambda c: (-float(c[0]) + 0.0, c[1]))
        for j, (score, _, idx) in enumerate(cand[:k_out]):
            D[q, j], I[q, j] = score, idx
    return D, I


def hybrid_fuse_ref(D, I, S, J, k_out):
    """Dense hits (D float32, I) + BM25 hits (S float64, J) -> fused (F float32, O, methods)."""
    D, I, S, J = np.asarray(D, np.float32), np.asarray(I), np.asarray(S, np.float64), np.asarray(J)
    nq = D.shape[0]
    F = np.full((nq, k_out), np.float32(-3.4028235e38), dtype=np.float32)
    O = np.full((nq, k_out), -1, dtype=np.int64)
    M = np.zeros((nq, k_out), dtype=np.uint8)
    for q in range(nq):
        dense = {int(i): float(d) for d, i in zip(D[q], I[q]) if i >= 0}
        bm = [(float(s), int(j)) for s, j in zip(S[q], J[q]) if j >= 0]
        bm_ids = {j for _, j in bm}
        fused = {}
        for i, d in dense.items():
            fused[i] = np.float32(np.float64(d) * (1.1 if i in bm_ids else 1.0))
        if bm:
            mn, mx = min(s for s, _ in bm), max(s for s, _ in bm)
            for s, j in bm:
                if j in dense:
                    continue
                norm = (s - mn) / (mx - mn) if mx > mn else (1.0 if mx > 0.0 else 0.0)
                fused[j] = np.float32(norm)
        order = sorted(fused.items(), key=lambda kv: (-float(kv[1]) + 0.0, kv[0]))[:k_out]
        for r, (idx, score) in enumerate(order):
            F[q, r], O[q, r] = score, idx
            M[q, r] = (1 if idx in dense else 0) | (2 if idx in bm_ids else 0)
    return F, O, M
