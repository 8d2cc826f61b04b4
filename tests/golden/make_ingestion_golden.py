"""Generate tests/golden/ingestion_outputs.json by running the REAL /root/reference/src/ingestion.py
(BM25Ingestor, VectorDBIngestor, unmodified) over the chunked reports of corpus.py.  `faiss` is replaced by a
row-collecting IndexFlatIP + the IxFI writer of corpus.py, `rank_bm25` by the numpy oracle, the embedding client by
stubs.ScriptedQwen; report iteration, chunk extraction, tokenisation, request batching (100 texts), array
conversion and file naming are the reference's own code.  Recorded per report: a digest of the rows written and the
pickled BM25 state.  tests/test_ingestion.py compares this repo's ingestors (GPU index objects) with it.

    python -m tests.golden.make_ingestion_golden
"""
import hashlib
import io
import json
import os
import pickle
import sys
import tempfile
import types
from contextlib import redirect_stdout
from pathlib import Path

import numpy as np

REF = "/root/reference"
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from tests.golden import corpus, make_golden, stubs  # noqa: E402

GOLDEN = os.path.join(ROOT, "tests", "golden", "ingestion_outputs.json")


class CountingQwen(stubs.ScriptedQwen):
    """Records the request sizes (the reference sends at most 100 texts per call, src/ingestion.py:71-72)."""

    def __init__(self):
        self.batches = []

    def get_embeddings(self, texts, model=None):
        self.batches.append(len(texts))
        return super().get_embeddings(texts, model)


def rows_digest(rows):
    rows = np.ascontiguousarray(rows, dtype="<f4")
    return {"shape": list(rows.shape), "sha256": hashlib.sha256(rows.tobytes()).hexdigest()}


def bm25_state(obj):
    """Comparable form of a BM25Okapi-like object's pickled attributes (rank_bm25's attribute set)."""
    return {"corpus_size": int(obj.corpus_size), "avgdl": float(obj.avgdl), "doc_len": [int(v) for v in obj.doc_len],
            "idf": sorted([k, float(v)] for k, v in obj.idf.items()), "average_idf": float(obj.average_idf),
            "k1": float(obj.k1), "b": float(obj.b), "epsilon": float(obj.epsilon),
            "doc_freqs_sha256": hashlib.sha256(json.dumps([sorted(d.items()) for d in obj.doc_freqs]).encode()).hexdigest()}


def big_report(n_chunks=230):
    """More than two embedding batches (230 chunks -> 100 + 100 + 30)."""
    rep = corpus.make_report("BIG230", corpus.COMPANY, 2018, n_pages=80, seed=901)
    rep["content"]["chunks"] = rep["content"]["chunks"][:n_chunks]
    assert len(rep["content"]["chunks"]) == n_chunks
    return rep


def build_reports_dir(tmp):
    paths = corpus.build(tmp)
    src = paths["bm25_documents_dir"]  # the clean directory (no broken / schema-less JSON: the ingestors have no guard)
    with open(Path(src) / "BIG230.json", "w", encoding="utf-8") as f:
        json.dump(big_report(), f)
    return Path(src)


def _install_stubs():
    make_golden.install_stubs()  # faiss.read_index, rank_bm25.BM25Okapi (oracle), src.api_requests, src.reranking

    class _Index:
        def __init__(self, d):
            self.d, self.parts = int(d), []

        def add(self, x):
            assert x.dtype == np.float32 and x.ndim == 2 and x.shape[1] == self.d
            self.parts.append(np.array(x))

        @property
        def rows(self):
            return np.concatenate(self.parts) if self.parts else np.zeros((0, self.d), np.float32)

    faiss = sys.modules["faiss"]
    faiss.IndexFlatIP = _Index
    faiss.write_index = lambda index, path: corpus.write_ixfi(path, index.rows)
    for name, attrs in (("dotenv", {"load_dotenv": lambda *a, **k: None}),
                        ("tenacity", {"retry": lambda **k: (lambda fn: fn), "wait_fixed": lambda s: None,
                                      "stop_after_attempt": lambda n: None})):
        try:
            __import__(name)
        except ImportError:
            mod = types.ModuleType(name)
            mod.__dict__.update(attrs)
            sys.modules[name] = mod


def main():
    _install_stubs()
    sys.path.insert(0, REF)
    import src.ingestion as ref

    assert ref.__file__.startswith(REF), ref.__file__
    out = {"vector": {}, "bm25": {}}
    with tempfile.TemporaryDirectory() as tmp, redirect_stdout(io.StringIO()):
        reports_dir = build_reports_dir(tmp)
        vec_out, bm_out = Path(tmp) / "out_vec", Path(tmp) / "out_bm25"
        ing = ref.VectorDBIngestor()
        ing.qwen = CountingQwen()
        ing.process_reports(reports_dir, vec_out)
        out["embedding_batches"] = sorted(ing.qwen.batches)
        for p in sorted(vec_out.glob("*.faiss")):
            out["vector"][p.stem] = rows_digest(stubs.read_ixfi(p).rows)
        ref.BM25Ingestor().process_reports(reports_dir, bm_out)
        for p in sorted(bm_out.glob("*.pkl")):
            with open(p, "rb") as f:
                out["bm25"][p.stem] = bm25_state(pickle.load(f))
        try:
            ing._get_embeddings("   ")
            out["empty_text_error"] = None
        except Exception as e:  # noqa: BLE001 - tenacity wraps the ValueError when it is the real package
            out["empty_text_error"] = type(e).__name__
    with open(GOLDEN, "w", encoding="utf-8") as f:
        json.dump(out, f, ensure_ascii=False, sort_keys=True, separators=(",", ":"))
    print("wrote", GOLDEN, os.path.getsize(GOLDEN), "bytes;", sorted(out["vector"]), out["embedding_batches"][-4:],
          out["empty_text_error"])


if __name__ == "__main__":
    main()
