"""CPU: pin the oracle with hand-derived known answers (the reference ships no golden vectors)."""
import math

import numpy as np
import pytest

import oracle
from oracle.bm25 import bm25_scores_csr_ref


def test_flat_ip_kat_with_tie_and_padding():
    xb = np.array([[1, 0, 0], [0, 1, 0], [1, 0, 0], [0.5, 0.5, 0]], dtype=np.float32)
    xq = np.array([[1, 0, 0], [0, 2, 0], [1, 1, 0]], dtype=np.float32)
    D, I = oracle.flat_ip_search_ref(xb, xq, 3)
    assert I.tolist() == [[0, 2, 3], [1, 3, 0], [0, 1, 2]]   # ties: ascending id
    assert D.tolist() == [[1.0, 1.0, 0.5], [2.0, 1.0, 0.0], [1.0, 1.0, 1.0]]
    D, I = oracle.flat_ip_search_ref(xb, xq[:1], 6)           # k > ntotal pads like faiss
    assert I.tolist() == [[0, 2, 3, 1, -1, -1]]
    assert D[0, 4] == np.float32(-3.4028235e38) and D.dtype == np.float32 and I.dtype == np.int64
    S64 = oracle.flat_ip_scores_ref(xb, xq, np.float64)
    S32 = oracle.flat_ip_scores_ref(xb, xq, np.float32)
    assert np.all(np.abs(S32 - S64) <= 1e-5 * np.abs(S64))
    oracle.assert_topk_equivalent(*oracle.flat_ip_search_ref(xb, xq, 2), S64, 2)


def test_bm25_kat_hand_computed():
    corpus = [["a", "b", "a"], ["b", "c"], ["c", "c", "d"]]
    bm = oracle.BM25OkapiRef(corpus)
    assert bm.corpus_size == 3 and bm.avgdl == 8 / 3 and bm.doc_len == [3, 2, 3]
    # idf = ln(N - n + 0.5) - ln(n + 0.5)
    idf_a = math.log(2.5) - math.log(1.5)      # in 1 doc
    idf_b = math.log(1.5) - math.log(2.5)      # in 2 docs -> negative
    avg = (idf_a + idf_b + idf_b + idf_a) / 4  # order of first appearance: a, b, c, d
    assert bm.average_idf == avg
    assert bm.idf["a"] == idf_a and bm.idf["d"] == idf_a
    assert bm.idf["b"] == 0.25 * avg and bm.idf["c"] == 0.25 * avg   # epsilon floor
    # score of doc 0 for query ["a"]: tf = 2, dl = 3
    w = 2 * 2.5 / (2 + 1.5 * (1 - 0.75 + 0.75 * 3 / (8 / 3)))
    s = bm.get_scores(["a"])
    assert s[0] == idf_a * w and s[1] == 0.0 and s[2] == 0.0
    # repeated token counted twice, unknown token ignored
    assert np.array_equal(bm.get_scores(["a", "a", "zz"]), 2 * s)
    # stable descending sort keeps ascending ids among ties, zero-score docs eligible
    assert oracle.bm25_topn_ref(s, 3) == [0, 1, 2]
    assert oracle.bm25_topn_ref(np.zeros(4), 2) == [0, 1]


def test_bm25_csr_form_matches_dict_form():
    rng = np.random.default_rng(0)
    docs = [[f"t{int(r)}" for r in rng.zipf(1.3, int(rng.integers(5, 40))) if r < 200] or ["t1"] for _ in range(300)]
    bm = oracle.BM25OkapiRef(docs)
    vocab = {w: i for i, w in enumerate(bm.idf)}
    terms, dids, tfs = [], [], []
    for di, f in enumerate(bm.doc_freqs):
        for w, c in f.items():
            terms.append(vocab[w]); dids.append(di); tfs.append(c)
    terms = np.array(terms); order = np.argsort(terms, kind="stable")
    indptr = np.zeros(len(vocab) + 1, np.int64)
    np.cumsum(np.bincount(terms, minlength=len(vocab)), out=indptr[1:])
    idf = np.array([bm.idf[w] for w in vocab])
    for q in (["t1", "t2"], ["t7", "t7", "t3"], ["t150"]):
        ids = [vocab[t] for t in q if t in vocab]
        got = bm25_scores_csr_ref(indptr, np.array(dids)[order], np.array(tfs)[order], np.array(bm.doc_len), bm.avgdl, idf, ids)
        assert np.array_equal(got, bm.get_scores(q))


def test_cosine_matrix_kat():
    x = np.array([[3, 4], [0, 2], [-3, -4]], dtype=np.float32)
    s = oracle.cosine_matrix_ref(x)
    assert np.allclose(s, [[1, 0.8, -1], [0.8, 1, -0.8], [-1, -0.8, 1]], atol=1e-6)
    st = oracle.cosine_stats_ref(s)
    assert abs(st["avg_similarity"] - (-1 / 3)) < 1e-6 and st["max_similarity"] == np.float32(0.8)
    assert abs(st["min_similarity"] + 1) < 1e-6 and abs(st["median_similarity"] + 0.8) < 1e-6


def test_oracle_cross_checked_with_sklearn_and_scipy():
    """Independent formulations from libraries that ARE installed here (neither faiss nor rank_bm25 is): sklearn's
    linear_kernel / cosine_similarity for the dense side, a scipy-sparse matrix form of the Okapi formula for BM25."""
    import scipy.sparse as sp
    from sklearn.feature_extraction.text import CountVectorizer
    from sklearn.metrics.pairwise import cosine_similarity, linear_kernel

    from rag_challenge_2_b200 import synth

    rng = np.random.default_rng(5)
    xb = rng.standard_normal((500, 64)).astype(np.float32)
    xq = rng.standard_normal((7, 64)).astype(np.float32)
    D, I = oracle.flat_ip_search_ref(xb, xq, 10)
    S = linear_kernel(xq.astype(np.float64), xb.astype(np.float64))
    order = np.argsort(-S, axis=1, kind="stable")[:, :10]
    assert np.array_equal(I, order)
    assert np.allclose(D, np.take_along_axis(S, order, 1), rtol=1e-5)

    x = rng.standard_normal((60, 32)).astype(np.float32) * rng.uniform(0.5, 2, (60, 1)).astype(np.float32)
    mine = oracle.cosine_matrix_ref(x)
    theirs = cosine_similarity(x.astype(np.float64))
    np.fill_diagonal(theirs, 1.0)
    assert np.allclose(mine, np.clip(theirs, -1, 1), atol=2e-6)

    docs = synth.token_corpus_np(300, vocab=400, seed=3)
    bm = oracle.BM25OkapiRef(docs)
    cv = CountVectorizer(analyzer=lambda d: d)
    tf = cv.fit_transform(docs).astype(np.float64).tocsc()            # [n_docs, n_terms]
    n = tf.shape[0]
    dl = np.asarray(tf.sum(axis=1)).ravel()
    df = np.diff(tf.indptr)
    idf = np.log(n - df + 0.5) - np.log(df + 0.5)
    idf[idf < 0] = 0.25 * idf.mean()                                  # epsilon floor with the mean over ALL terms' raw idf
    k1, b = 1.5, 0.75
    norm = k1 * (1 - b + b * dl / dl.mean())
    for q in synth.token_queries_np(6, vocab=400, seed=4):
        toks = [f"t{r}" for r in q]
        want = np.zeros(n)
        for t in toks:
            j = cv.vocabulary_.get(t)
            if j is None:
                continue
            col = np.asarray(tf[:, j].todense()).ravel()
            want += idf[j] * col * (k1 + 1) / (col + norm)
        assert np.allclose(bm.get_scores(toks), want, rtol=1e-12, atol=1e-12)
    assert sp.issparse(tf)


# ---- vectors PUBLISHED by the two third-party packages themselves --------------------------------------------
# rank_bm25's README ("Ranking of documents"): the only numeric output rank-bm25 0.2.2 documents.
RANK_BM25_README_CORPUS = ["Hello there good man!", "It is quite windy in London", "How is the weather today?"]
RANK_BM25_README_SCORES = [0.0, 0.93729472, 0.0]

# faiss wiki, "Getting started" (IndexFlatL2, d = 64, nb = 100000, nq = 10000, np.random.seed(1234), k = 4): the printed
# sanity check (the first five database rows searched against the index) and the first / last five query results.
FAISS_WIKI_SANITY_I = [[0, 393, 363, 78], [1, 555, 277, 364], [2, 304, 101, 13], [3, 173, 18, 182], [4, 288, 370, 531]]
FAISS_WIKI_SANITY_D = [[0.0, 7.17517328, 7.2076292, 7.25116253], [0.0, 6.32356453, 6.6845808, 6.79994535],
                       [0.0, 5.79640865, 6.39173603, 7.28151226], [0.0, 7.27790546, 7.52798653, 7.66284657],
                       [0.0, 6.76380348, 7.29512024, 7.36881447]]
FAISS_WIKI_FIRST_I = [[381, 207, 210, 477], [526, 911, 142, 72], [838, 527, 1290, 425], [196, 184, 164, 359], [526, 377, 120, 425]]
FAISS_WIKI_LAST_I = [[9900, 10500, 9309, 9831], [11055, 10895, 10812, 11321], [11353, 11103, 10164, 9787],
                     [10571, 10664, 10632, 9638], [9628, 9554, 10036, 9582]]


def faiss_wiki_dataset():
    """The tutorial's data, bit for bit (numpy's legacy MT19937 stream is frozen), and its IndexFlatL2 search restated as an
    inner-product search: ||q - x||^2 = ||q||^2 - (2 q.x - ||x||^2), so ranking by L2 distance ascending is ranking by the
    inner product of [q, 1] with [2x, -||x||^2] descending — what IndexFlatIP (and this repo's kernels) compute."""
    d, nb, nq = 64, 100000, 10000
    np.random.seed(1234)
    xb = np.random.random((nb, d)).astype("float32")
    xb[:, 0] += np.arange(nb) / 1000.0
    xq = np.random.random((nq, d)).astype("float32")
    xq[:, 0] += np.arange(nq) / 1000.0
    rows = np.concatenate([2 * xb, -np.sum(xb.astype(np.float64) ** 2, axis=1, keepdims=True).astype(np.float32)], axis=1)
    return xb, xq, np.ascontiguousarray(rows)


def faiss_wiki_queries(x):
    return np.ascontiguousarray(np.concatenate([x, np.ones((len(x), 1), np.float32)], axis=1))


def test_rank_bm25_published_readme_vector():
    """The README's example: query "windy London" -> array([0., 0.93729472, 0.]) and get_top_n(..., n=1) -> doc 1."""
    corpus = [doc.split(" ") for doc in RANK_BM25_README_CORPUS]
    bm = oracle.BM25OkapiRef(corpus)
    scores = bm.get_scores("windy London".split(" "))
    assert np.round(scores, 8).tolist() == RANK_BM25_README_SCORES
    assert oracle.bm25_topn_ref(scores, 1) == [1]


def test_faiss_published_getting_started_vectors():
    """The oracle's search (scores, selection, order) reproduces every id the faiss tutorial prints, and its distances."""
    xb, xq, rows = faiss_wiki_dataset()
    for x, want_i, want_d in ((xb[:5], FAISS_WIKI_SANITY_I, FAISS_WIKI_SANITY_D), (xq[:5], FAISS_WIKI_FIRST_I, None),
                              (xq[-5:], FAISS_WIKI_LAST_I, None)):
        D, I = oracle.flat_ip_search_ref(rows, faiss_wiki_queries(x), 4)
        assert I.tolist() == want_i
        if want_d is not None:
            l2 = np.sum(x.astype(np.float64) ** 2, axis=1, keepdims=True) - D
            assert np.allclose(l2, want_d, atol=2e-4)


def test_oracle_against_the_real_packages_when_installed():
    """faiss-cpu / rank-bm25 are not installable in this image; wherever they are, this pins the
    restatement to them directly."""
    from rag_challenge_2_b200 import synth

    checked = 0
    try:
        import faiss
    except ImportError:
        faiss = None
    if faiss is not None:
        rng = np.random.default_rng(11)
        xb = rng.standard_normal((2000, 64)).astype(np.float32)
        xb[100] = xb[7]                                               # an exact tie
        xq = np.concatenate([rng.standard_normal((5, 64)).astype(np.float32), xb[7:8]])
        index = faiss.IndexFlatIP(64)
        index.add(xb)
        for k in (1, 10, 30):
            D, I = index.search(xq, k)
            oracle.assert_topk_equivalent(D, I, oracle.flat_ip_scores_ref(xb, xq, np.float64), k, what=f"faiss k={k}")
            Do, Io = oracle.flat_ip_search_ref(xb, xq, k)
            assert np.array_equal(I[-1, :2] if k > 1 else I[-1, :1], Io[-1, :2] if k > 1 else Io[-1, :1])   # tie order
        small = faiss.IndexFlatIP(64)
        small.add(xb[:2])
        D, I = small.search(xq[:1], 5)
        assert list(I[0, 2:]) == [-1, -1, -1] and np.all(D[0, 2:] == oracle.flat_ip.NEG)   # padding
        checked += 1
    try:
        import rank_bm25
    except ImportError:
        rank_bm25 = None
    if rank_bm25 is not None:
        docs = synth.token_corpus_np(400, vocab=300, seed=12)
        real, mine = rank_bm25.BM25Okapi(docs), oracle.BM25OkapiRef(docs)
        assert real.avgdl == mine.avgdl and real.idf == mine.idf and real.average_idf == mine.average_idf
        for q in synth.token_queries_np(8, vocab=300, seed=13):
            toks = [f"t{r}" for r in q] + ["unknown-token"]
            assert np.array_equal(real.get_scores(toks), mine.get_scores(toks))
        checked += 1
    if not checked:
        pytest.skip("neither faiss nor rank_bm25 is installed here")
