"""Oracle for rank_bm25.BM25Okapi (rank-bm25 0.2.2), restated from its published source
(rank_bm25.py: class BM25.__init__/_initialize, class BM25Okapi.__init__/_calc_idf/get_scores).

Call sites in the reference: src/ingestion.py:19-22 (build, tokeniser = str.split),
src/retrieval.py:261-265 (get_scores + full stable sort), :270 (round(score, 4)).

Everything is float64, exactly as numpy computes it there; the operation order inside get_scores is
kept (num = tf*(k1+1); den = tf + k1*(1 - b + b*dl/avgdl); score += idf*(num/den)) so that an
implementation using the same order can be compared bit-for-bit.
"""
import math

import numpy as np


class BM25OkapiRef:
    def __init__(self, corpus, k1=1.5, b=0.75, epsilon=0.25):
        self.k1, self.b, self.epsilon = k1, b, epsilon
        self.corpus_size = 0
        self.doc_len = []
        self.doc_freqs = []
        self.idf = {}
        nd = {}
        num_tokens = 0
        for document in corpus:
            self.doc_len.append(len(document))
            num_tokens += len(document)
            freqs = {}
            for word in document:
                freqs[word] = freqs.get(word, 0) + 1
            self.doc_freqs.append(freqs)
            for word in freqs:
                nd[word] = nd.get(word, 0) + 1
            self.corpus_size += 1
        self.avgdl = num_tokens / self.corpus_size  # ZeroDivisionError on an empty corpus, as upstream
        self.nd = nd
        idf_sum = 0
        negative = []
        for word, freq in nd.items():
            idf = math.log(self.corpus_size - freq + 0.5) - math.log(freq + 0.5)
            self.idf[word] = idf
            idf_sum += idf
            if idf < 0:
                negative.append(word)
        self.average_idf = idf_sum / len(self.idf)
        eps = self.epsilon * self.average_idf
        for word in negative:
            self.idf[word] = eps

    def get_scores(self, query):
        score = np.zeros(self.corpus_size)
        doc_len = np.array(self.doc_len)
        for q in query:
            q_freq = np.array([(doc.get(q) or 0) for doc in self.doc_freqs])
            score += (self.idf.get(q) or 0) * (
                q_freq * (self.k1 + 1) / (q_freq + self.k1 * (1 - self.b + self.b * doc_len / self.avgdl))
            )
        return score


def bm25_topn_ref(scores, top_n):
    """src/retrieval.py:264-265 — full stable sort, descending; equal scores keep ascending index."""
    n = min(top_n, len(scores))
    return sorted(range(len(scores)), key=lambda i: scores[i], reverse=True)[:n]


def bm25_scores_csr_ref(indptr, doc_ids, tf, doc_len, avgdl, idf, term_ids, k1=1.5, b=0.75):
    """Same arithmetic as BM25OkapiRef.get_scores over term-major CSR postings (vectorised; used for
    corpora too large for the dict-of-dicts form).  Validated against BM25OkapiRef in the tests."""
    doc_len = np.asarray(doc_len, dtype=np.int64)
    score = np.zeros(doc_len.shape[0])
    for t in term_ids:
        lo, hi = int(indptr[t]), int(indptr[t + 1])
        d = doc_ids[lo:hi]
        f = tf[lo:hi].astype(np.int64)
        score[d] += idf[t] * (f * (k1 + 1) / (f + k1 * (1 - b + b * doc_len[d] / avgdl)))
    return score


def token_corpus_arrays_np(n_docs, vocab=50_000, seed=4242, mean_len=120.0, sigma=0.5):
    """The synthetic token corpus of SURVEY 8d (Zipf(1.0) ranks, lognormal lengths clipped to [8, 512]) as flat
    arrays — (ranks int32[total], lens int64[n_docs]) — generated vectorised so that a million documents take
    seconds, not minutes.  Same law as rag_challenge_2_b200.synth.token_corpus_np, not the same random stream."""
    rng = np.random.default_rng(seed)
    w = 1.0 / np.arange(1, vocab + 1, dtype=np.float64)
    cdf = np.cumsum(w / w.sum())
    lens = np.clip(np.rint(rng.lognormal(math.log(mean_len), sigma, n_docs)), 8, 512).astype(np.int64)
    total = int(lens.sum())
    ranks = np.empty(total, dtype=np.int32)
    step = 1 << 24
    for lo in range(0, total, step):
        hi = min(total, lo + step)
        ranks[lo:hi] = np.minimum(np.searchsorted(cdf, rng.random(hi - lo), side="left"), vocab - 1)
    return ranks, lens


def bm25_ref_from_arrays(ranks, lens, vocab, k1=1.5, b=0.75, epsilon=0.25):
    """A BM25OkapiRef — the rank_bm25 structure itself: one {token: tf} dict per document, idf dict, doc_len list —
    built from the flat arrays above without the per-token Python loop of __init__ (that loop is not what the
    bench times; get_scores, the per-query loop over the dictionaries, is).  Tokens are the strings "t<rank>";
    the statistics are computed exactly as __init__ computes them (first-seen term order for the idf average)."""
    n_docs = int(lens.shape[0])
    names = [f"t{r}" for r in range(vocab)]
    doc_of = np.repeat(np.arange(n_docs, dtype=np.int64), lens)
    key = doc_of * vocab + ranks
    # first position of every term: fancy assignment keeps the LAST value written to a repeated index, so writing the
    # positions in descending order leaves the first occurrence
    first_seen = np.full(vocab, np.iinfo(np.int64).max, dtype=np.int64)
    first_seen[ranks[::-1]] = np.arange(len(ranks) - 1, -1, -1, dtype=np.int64)
    del doc_of
    key.sort()
    new = np.empty(key.shape[0], dtype=bool)
    new[0] = True
    np.not_equal(key[1:], key[:-1], out=new[1:])
    idx = np.flatnonzero(new)
    tfs = np.diff(np.r_[idx, key.shape[0]])
    uniq = key[idx]
    del key, new, idx
    docs = uniq // vocab
    terms = (uniq - docs * vocab).astype(np.int64)
    del uniq
    bounds = np.zeros(n_docs + 1, dtype=np.int64)
    np.cumsum(np.bincount(docs, minlength=n_docs), out=bounds[1:])
    self = BM25OkapiRef.__new__(BM25OkapiRef)
    self.k1, self.b, self.epsilon = k1, b, epsilon
    self.corpus_size = n_docs
    self.doc_len = lens.tolist()
    self.avgdl = int(lens.sum()) / n_docs
    tl, fl, bl = terms.tolist(), tfs.tolist(), bounds.tolist()
    getname = names.__getitem__
    self.doc_freqs = [dict(zip(map(getname, tl[bl[i]:bl[i + 1]]), fl[bl[i]:bl[i + 1]])) for i in range(n_docs)]
    df = np.bincount(terms, minlength=vocab)
    present = np.flatnonzero(df > 0)
    present = present[np.argsort(first_seen[present], kind="stable")]   # dict insertion order of rank_bm25's nd
    self.nd = {names[t]: int(df[t]) for t in present.tolist()}
    self.idf = {}
    idf_sum = 0
    negative = []
    for word, freq in self.nd.items():
        idf = math.log(self.corpus_size - freq + 0.5) - math.log(freq + 0.5)
        self.idf[word] = idf
        idf_sum += idf
        if idf < 0:
            negative.append(word)
    self.average_idf = idf_sum / len(self.idf)
    eps = self.epsilon * self.average_idf
    for word in negative:
        self.idf[word] = eps
    return self
