"""GPU parity against the only numeric vectors the two third-party packages PUBLISH: rank_bm25's README example and the
faiss wiki's "Getting started" output (tests/test_oracle_kat.py holds the constants and pins the oracle to them on the CPU).
Everything goes through the C ABI (FlatIPIndex / BM25Okapi -> libb2r.so)."""
import numpy as np
import pytest

from tests.test_oracle_kat import (FAISS_WIKI_FIRST_I, FAISS_WIKI_LAST_I, FAISS_WIKI_SANITY_D, FAISS_WIKI_SANITY_I,
                                   RANK_BM25_README_CORPUS, RANK_BM25_README_SCORES, faiss_wiki_dataset, faiss_wiki_queries)

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("bits", [64, 32])
def test_rank_bm25_readme_vector(b2r, bits):
    from rag_challenge_2_b200.bm25 import BM25Okapi

    bm = BM25Okapi([doc.split(" ") for doc in RANK_BM25_README_CORPUS], weight_bits=bits)
    scores = bm.get_scores("windy London".split(" "))
    S, I = bm.topk(["windy London".split(" ")], 1)
    assert I.tolist() == [[1]] and S[0, 0] == scores[1]
    if bits == 64:       # the default layout: every printed digit
        assert np.round(scores, 8).tolist() == RANK_BM25_README_SCORES
    else:                # float32 weights (explicit opt-in): 1e-6 relative, north_star's bar is 1e-5
        assert scores[0] == 0.0 and scores[2] == 0.0 and abs(scores[1] - RANK_BM25_README_SCORES[1]) <= 1e-6 * scores[1]


@pytest.mark.parametrize("pad_to", [65, 128])
def test_faiss_getting_started_vectors(b2r, pad_to):
    """d = 65 runs the generic-width scan, the zero-padded d = 128 copy the bulk-copy (TMA) scan; 10 queries at once also
    cross the hand-over to the tensor-core batch path."""
    xb, xq, rows = faiss_wiki_dataset()
    if pad_to > rows.shape[1]:
        rows = np.ascontiguousarray(np.pad(rows, ((0, 0), (0, pad_to - rows.shape[1]))))
    index = b2r.FlatIPIndex(pad_to)
    index.add(rows)

    def q(x):
        y = faiss_wiki_queries(x)
        return np.ascontiguousarray(np.pad(y, ((0, 0), (0, pad_to - y.shape[1]))))

    D, I = index.search(q(xb[:5]), 4)
    assert I.tolist() == FAISS_WIKI_SANITY_I
    l2 = np.sum(xb[:5].astype(np.float64) ** 2, axis=1, keepdims=True) - D
    assert np.allclose(l2, FAISS_WIKI_SANITY_D, atol=2e-4)
    assert index.search(q(xq[:5]), 4)[1].tolist() == FAISS_WIKI_FIRST_I
    assert index.search(q(xq[-5:]), 4)[1].tolist() == FAISS_WIKI_LAST_I
    for i in range(5):                                   # one query at a time: the single-query scan
        assert index.search(q(xq[i: i + 1]), 4)[1].tolist() == [FAISS_WIKI_FIRST_I[i]]
    both = index.search(q(np.concatenate([xq[:5], xq[-5:]])), 4)[1].tolist()
    assert both == FAISS_WIKI_FIRST_I + FAISS_WIKI_LAST_I
