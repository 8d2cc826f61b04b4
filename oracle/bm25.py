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
