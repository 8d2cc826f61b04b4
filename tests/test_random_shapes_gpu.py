"""Seeded sweep over random shapes through the public classes (C ABI underneath) against the oracle.

compute-sanitizer is not available on the GPU pool, so memory safety rests on parity at awkward shapes: index sizes
around the tile / block / warp boundaries, k from 1 to past the index size, widths on and off the bulk-copy path,
batches that cross the SIMT / tensor-core dispatch, BM25 corpora with empty documents, unknown / repeated query tokens
and queries longer than a warp, and the hybrid step on top of both.  Every case is reproducible from its seed."""
import numpy as np
import pytest

import oracle

pytestmark = pytest.mark.gpu

WIDTHS = [1, 2, 5, 31, 33, 64, 100, 127, 128, 129, 200, 256, 384, 500, 512, 768, 1000, 1024, 1536, 2048]
SIZES = [1, 2, 3, 31, 32, 33, 63, 64, 65, 127, 128, 129, 255, 257, 1000, 4095, 4096, 4097, 9471, 9473, 20000]


@pytest.mark.parametrize("seed", range(120))
def test_flat_search_random_shapes(b2r, seed):
    rng = np.random.default_rng(1000 + seed)
    d = int(rng.choice(WIDTHS))
    n = int(rng.choice(SIZES)) if rng.random() < 0.7 else int(rng.integers(1, 30000))
    if n * d > 24_000_000:
        n = 24_000_000 // d
    nq = int(rng.choice([1, 1, 2, 3, 4, 5, 7, 8, 9, 17, 33, 64, 130, 257]))
    k = int(rng.choice([1, 2, 10, 30, 31, 50, 64, 100, 128, 255, 256, 257, 512, 1024]))
    xb = rng.standard_normal((n, d), dtype=np.float32)
    if rng.random() < 0.5:
        xb /= np.maximum(np.linalg.norm(xb, axis=1, keepdims=True), 1e-12)
    if n > 4 and rng.random() < 0.3:            # duplicate rows: ties must resolve to the lower id
        src = rng.integers(0, n, max(1, n // 10))
        dst = rng.integers(0, n, src.size)
        xb[dst] = xb[src]
    xq = rng.standard_normal((nq, d), dtype=np.float32)
    if n > 1 and rng.random() < 0.5:
        xq[::2] = xb[rng.integers(0, n, xq[::2].shape[0])]
    index = b2r.FlatIPIndex(d)
    cut = int(rng.integers(0, n + 1))            # two add() calls: the second starts off a tile boundary
    index.add(xb[:cut])
    index.add(xb[cut:])
    assert index.ntotal == n
    D, I = index.search(xq, k)
    S = oracle.flat_ip_scores_ref(xb, xq, np.float64)
    # 1e-5 relative (north_star); the absolute floor scales with |q| |x| — unnormalised rows can cancel to a score near 0
    scale = float(np.linalg.norm(xb, axis=1).max() * np.linalg.norm(xq, axis=1).max())
    oracle.assert_topk_equivalent(D, I, S, k, rtol=1e-5, atol=1e-6 * max(1.0, scale),
                                  what=f"seed={seed} n={n} d={d} nq={nq} k={k}")
    if k > n:                                    # faiss padding: id -1, score -FLT_MAX
        assert np.all(I[:, n:] == -1) and np.all(D[:, n:] == np.float32(-3.4028235e38))
    index.close()


def _random_corpus(rng, n_docs, vocab):
    lens = rng.integers(0, 40, n_docs)
    lens[rng.random(n_docs) < 0.05] = 0          # empty documents
    if n_docs > 3:
        lens[int(rng.integers(0, n_docs))] = 700  # one long document
    zipf = 1.0 / np.arange(1, vocab + 1) ** 1.1
    zipf /= zipf.sum()
    return [[f"w{t}" for t in rng.choice(vocab, int(m), p=zipf)] for m in lens]


@pytest.mark.parametrize("seed", range(60))
def test_bm25_random_corpora(b2r, seed):
    from rag_challenge_2_b200.bm25 import BM25Okapi

    rng = np.random.default_rng(2000 + seed)
    n_docs = int(rng.choice([1, 2, 3, 31, 33, 255, 257, 1000, 4095, 4096, 4097, 8193, 12000]))
    vocab = int(rng.choice([1, 2, 7, 50, 400, 3000]))
    docs = _random_corpus(rng, n_docs, vocab)
    if not any(docs):
        docs[0] = ["w0"]
    ref = oracle.BM25OkapiRef(docs)
    idx = BM25Okapi(docs)                        # default: float64 weights, bit-exact
    queries = []
    for _ in range(int(rng.choice([1, 2, 5, 9]))):
        m = int(rng.choice([0, 1, 2, 3, 8, 33, 70]))
        q = [f"w{t}" for t in rng.integers(0, vocab + 3, m)]   # ids >= vocab: unknown tokens
        if m > 1 and rng.random() < 0.5:
            q[-1] = q[0]                         # a repeated token counts twice
        queries.append(q)
    k = int(rng.choice([1, 3, 30, 64, 200]))
    S, I = idx.topk(queries, k)
    for qi, q in enumerate(queries):
        want = ref.get_scores(q)
        assert np.array_equal(idx.get_scores(q), want), (seed, qi)
        top = oracle.bm25_topn_ref(want, k)
        m = len(top)
        assert np.array_equal(I[qi, :m], top), (seed, qi, n_docs, vocab, k)
        assert np.array_equal(S[qi, :m], want[top]), (seed, qi)
        assert np.all(I[qi, m:] == -1)


@pytest.mark.parametrize("seed", range(30))
def test_hybrid_step_random_shapes(b2r, seed):
    from rag_challenge_2_b200 import hybrid
    from rag_challenge_2_b200.bm25 import BM25Okapi

    rng = np.random.default_rng(3000 + seed)
    n = int(rng.choice([1, 5, 40, 129, 1000, 4097, 9000]))
    d = int(rng.choice([64, 128, 1024]))
    k = int(rng.choice([1, 5, 30, 100, 256]))
    vocab = int(rng.choice([3, 40, 800]))
    docs = _random_corpus(rng, n, vocab)
    if not any(docs):
        docs[0] = ["w0"]
    xb = rng.standard_normal((n, d), dtype=np.float32)
    xb /= np.linalg.norm(xb, axis=1, keepdims=True)
    flat = b2r.FlatIPIndex(d)
    flat.add(xb)
    bm = BM25Okapi(docs)
    ref = oracle.BM25OkapiRef(docs)
    hy = hybrid.HybridIndex(flat, bm, 0, world_size=1)
    for nq in (1, 2, 3):
        xq = rng.standard_normal((nq, d), dtype=np.float32)
        toks = [[f"w{t}" for t in rng.integers(0, vocab + 2, int(rng.choice([0, 1, 4, 12])))] for _ in range(nq)]
        tids = [bm._term_ids(q).tolist() for q in toks]
        F, O = hy.search_hybrid(xq, tids, k)
        # the two legs against the oracle (dense within 1e-5, BM25 bit for bit) ...
        Dg, Ig = flat.search(xq, k)
        oracle.assert_topk_equivalent(Dg, Ig, oracle.flat_ip_scores_ref(xb, xq, np.float64), k, rtol=1e-5,
                                      what=f"seed={seed} n={n} d={d} nq={nq} k={k}")
        Sg, Jg = bm.topk(toks, k)
        for qi, q in enumerate(toks):
            sc = ref.get_scores(q)
            top = oracle.bm25_topn_ref(sc, k)
            assert np.array_equal(Jg[qi, :len(top)], top) and np.array_equal(Sg[qi, :len(top)], sc[top]), (seed, qi)
            assert np.all(Jg[qi, len(top):] == -1)
        # ... and the step == the oracle's fusion of exactly those legs (ids and float32 scores identical)
        Fr, Or, _ = oracle.hybrid_fuse_ref(Dg, Ig, Sg, Jg, k)
        assert np.array_equal(O, Or), (seed, n, d, k, nq)
        assert np.array_equal(F, Fr), (seed, n, d, k, nq)


@pytest.mark.parametrize("seed", range(24))
def test_cosine_stats_random_shapes(b2r, seed):
    """K5 (CTA-pair kernel where the width allows, the single-CTA kernel otherwise) against the numpy restatement of
    generate_similarity_matrix.py: sizes around the 128 / 256-row tile edges, widths on and off the tensor-core path."""
    import torch

    from rag_challenge_2_b200 import cosine

    rng = np.random.default_rng(4000 + seed)
    n = int(rng.choice([2, 3, 127, 128, 129, 255, 256, 257, 511, 513, 1000, 1025, 2049, 3000]))
    d = int(rng.choice([8, 64, 96, 100, 128, 256, 500, 1024, 1536]))
    x = rng.standard_normal((n, d)).astype(np.float32) + 0.5 * rng.standard_normal((1, d)).astype(np.float32)
    x *= rng.uniform(0.25, 4.0, (n, 1)).astype(np.float32)
    if n > 8:
        x[5] = 2.0 * x[1]                        # a pair that must clip at exactly 1.0
    got = cosine.cosine_stats(torch.from_numpy(x).cuda())
    want = oracle.cosine_stats_ref(oracle.cosine_matrix_ref(x))
    assert got["count"] == n * (n - 1)
    for key in ("avg_similarity", "std_similarity", "max_similarity", "min_similarity"):
        assert abs(got[key] - want[key]) <= 1e-5 * abs(want[key]) + 2e-6, (seed, n, d, key, got[key], want[key])
    med_tol = 2e-6 if n <= 1024 else 1e-4
    assert abs(got["median_similarity"] - want["median_similarity"]) <= med_tol, (seed, n, d)
