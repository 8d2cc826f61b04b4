"""Deterministic stand-ins shared by make_golden.py (which runs the REAL reference code with them)
and by the tests (which run this repo's retrievers with them): an embedder / LLM with the surface
of src.api_requests.BaseQwenProcessor, a reranker with the surface of src.reranking.LLMReranker, and
oracle-backed index objects with the faiss / rank_bm25 surface.  Test infrastructure only."""
import hashlib
import re
import struct

import numpy as np

import oracle

DIM = 1024


def _word_vec(word):
    seed = int(hashlib.sha1(word.encode("utf-8")).hexdigest()[:16], 16)
    return np.random.default_rng(seed).standard_normal(DIM).astype(np.float32)


_WORD_CACHE = {}


def embed_text(text):
    """Bag-of-words embedding: normalised sum of per-word hash vectors, so texts sharing words are
    similar (graded cosine scores instead of pure noise)."""
    words = re.findall(r"[A-Za-z0-9]+|[一-鿿]", text.lower()) or ["<empty>"]
    acc = np.zeros(DIM, dtype=np.float32)
    for w in words:
        v = _WORD_CACHE.get(w)
        if v is None:
            v = _WORD_CACHE[w] = _word_vec(w)
        acc += v
    return acc / np.linalg.norm(acc)


class ScriptedQwen:
    """BaseQwenProcessor surface: get_embeddings(texts) and send_message(...) (scripted LLM)."""

    def get_embeddings(self, texts, model=None):
        if isinstance(texts, str):
            texts = [texts]
        return [{"embedding": embed_text(t).tolist(), "text": t} for t in texts]

    def send_message(self, model=None, system_content="", human_content="", is_structured=False, **kw):
        m = re.search(r"(?:question|Original question): (.*?)(?:\n|$)", human_content)
        question = m.group(1).strip() if m else human_content[-80:]
        words = question.split()
        if "markdown" in human_content and "table" in human_content and "Original question" not in human_content:
            # HyDE: a hypothetical table built from the question's words
            return "| item | amount |\n|---|---|\n| " + " | ".join(words[:6]) + " |\n| revenue margin cash | 2024 |"
        # multi-query: the three methods are told apart by their instruction text
        if "子问题" in human_content:
            return f"<{' '.join(words[:4])} revenue> <SKIP> <{' '.join(words[-4:])} margin>"
        if "变体" in human_content or "情景" in human_content:
            return "<SKIP>"
        return f"<{question} definition gross margin ratio> < > <{question}>"


class ScriptedReranker:
    """LLMReranker surface (src/reranking.py:286-379): relevance from a hash of the text."""

    def __init__(self, *a, **kw):
        self.calls = 0

    def rerank_documents(self, query, documents, documents_batch_size=4, llm_weight=0.7, progress_callback=None):
        self.calls += 1
        out = []
        for doc in documents:
            d = doc.copy()
            h = int(hashlib.sha1((query + d["text"]).encode("utf-8")).hexdigest()[:8], 16)
            d["relevance_score"] = round((h % 101) / 100.0, 2)
            d["reasoning"] = "scripted"
            d["combined_score"] = round(d["relevance_score"] * doc.get("vector_similarity", doc.get("distance", 0.0)), 4)
            out.append(d)
        out.sort(key=lambda x: x["combined_score"], reverse=True)
        return out

    def get_stats(self):
        return {"total_requests": 0, "success_rate": 0.0, "scripted": True}


class OracleFlatIndex:
    """faiss.IndexFlatIP surface on top of the numpy oracle."""

    def __init__(self, rows):
        self.rows = np.ascontiguousarray(rows, dtype=np.float32)
        self.ntotal, self.d = self.rows.shape

    def search(self, x, k):
        # one query at a time: numpy's fp32 GEMV and GEMM differ in the last bit, and the golden outputs were
        # produced by the reference, which always searches a single query (src/retrieval.py:1140)
        x = np.ascontiguousarray(x, dtype=np.float32).reshape(-1, self.d)
        parts = [oracle.flat_ip_search_ref(self.rows, x[i: i + 1], k) for i in range(x.shape[0])]
        return np.concatenate([p[0] for p in parts], axis=0), np.concatenate([p[1] for p in parts], axis=0)

    def reconstruct(self, i):
        i = int(i)
        if not 0 <= i < self.ntotal:
            raise RuntimeError(f"id {i} out of range")
        return self.rows[i].copy()

    def reconstruct_batch(self, ids):
        ids = np.asarray(ids, dtype=np.int64)
        if ids.size and (ids.min() < 0 or ids.max() >= self.ntotal):  # faiss raises; numpy would wrap -1 around
            raise RuntimeError("id out of range")
        return self.rows[ids].copy()


def read_ixfi(path):
    """Independent parser of the IndexFlatIP file layout (see flat_index.py for the format notes)."""
    with open(str(path), "rb") as f:
        fourcc, d, ntotal = struct.unpack("<4siq", f.read(16))
        f.read(8 + 8 + 1 + 4)
        (n_floats,) = struct.unpack("<Q", f.read(8))
        assert fourcc == b"IxFI" and n_floats == d * ntotal
        rows = np.frombuffer(f.read(4 * n_floats), dtype="<f4").reshape(ntotal, d)
    return OracleFlatIndex(rows)


class OracleBM25(oracle.BM25OkapiRef):
    """rank_bm25.BM25Okapi surface (oracle) plus the fused topk() this repo's retriever calls."""

    def topk(self, queries, k):
        S = np.full((len(queries), k), -1.7976931348623157e308)
        I = np.full((len(queries), k), -1, dtype=np.int64)
        for qi, q in enumerate(queries):
            scores = self.get_scores(q)
            top = oracle.bm25_topn_ref(scores, k)
            I[qi, :len(top)] = top
            S[qi, :len(top)] = scores[top]
        return S, I
