"""Generate tests/golden/reference_outputs.json by running the REAL reference code
(/root/reference/src/retrieval.py, unmodified) on the miniature corpus of corpus.py.

The reference cannot be imported as-is in this container: faiss, rank_bm25 and the network
clients are missing.  They are replaced here — and only here — by stand-ins: `faiss` and
`rank_bm25` by the numpy oracle (oracle/), `src.api_requests` / `src.reranking` by the scripted
services in stubs.py.  Everything else (routing, query expansion protocol, per-hit record building,
aggregation, score fusion, sorting, truncation, BM25 retriever glue) is the reference's own code.

Run from the repo root in the authoring container:   python -m tests.golden.make_golden
"""
import json
import os
import sys
import tempfile
import types
from pathlib import Path

REF = "/root/reference"
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from tests.golden import cases, corpus, stubs  # noqa: E402


def install_stubs():
    faiss = types.ModuleType("faiss")
    faiss.read_index = lambda path: stubs.read_ixfi(path)
    faiss.IndexFlatIP = None
    sys.modules["faiss"] = faiss

    rank_bm25 = types.ModuleType("rank_bm25")

    class BM25Okapi(stubs.OracleBM25):
        pass

    BM25Okapi.__module__ = "rank_bm25"
    BM25Okapi.__qualname__ = "BM25Okapi"
    rank_bm25.BM25Okapi = BM25Okapi
    sys.modules["rank_bm25"] = rank_bm25

    api = types.ModuleType("src.api_requests")
    api.BaseQwenProcessor = stubs.ScriptedQwen
    api.BaseGeminiProcessor = object
    sys.modules["src.api_requests"] = api

    rer = types.ModuleType("src.reranking")
    rer.LLMReranker = stubs.ScriptedReranker
    sys.modules["src.reranking"] = rer


def main():
    install_stubs()
    sys.path.insert(0, REF)
    import src.retrieval as ref  # the reference's own module

    assert ref.__file__.startswith(REF), ref.__file__
    out = {"vector": {}, "hybrid": {}, "bm25": {}, "route": [], "years": [], "errors": {}}
    with tempfile.TemporaryDirectory() as tmp:
        paths = corpus.build(tmp)
        common = dict(vector_db_dir=paths["vector_db_dir"], documents_dir=paths["documents_dir"],
                      subset_path=paths["subset_path"], parallel_workers=1)  # 1 worker: deterministic completion order
        for name, kw in cases.VECTOR_CASES.items():
            kw = dict(kw)
            init = kw.pop("_init", {})
            r = ref.VectorRetriever(**common, **init)
            out["vector"][name] = cases.normalise(r.retrieve_by_company_name(**kw))
        for name, kw in cases.HYBRID_CASES.items():
            r = ref.HybridRetriever(**common)
            out["hybrid"][name] = cases.normalise(r.retrieve_by_company_name(**kw))
        r = ref.VectorRetriever(**common)
        out["retrieve_all"] = cases.normalise(r.retrieve_all(corpus.OTHER))
        out["loaded_reports"] = sorted(db["name"] for db in r.all_dbs)
        out["year_lookup"] = r.year_lookup
        for name, fn in (("vector_unknown_company", lambda: r.retrieve_by_company_name("nobody", "q", use_hyde=False, use_multi_query=False)),
                         ("retrieve_all_unknown_company", lambda: r.retrieve_all("nobody"))):
            try:
                fn()
                out["errors"][name] = None
            except Exception as e:  # noqa: BLE001
                out["errors"][name] = [type(e).__name__, str(e)]
        for kw in cases.ROUTE_CASES:
            got = ref.route_reports_by_time(all_reports=r.all_dbs, **kw)
            out["route"].append(sorted(rep["name"] for rep in got))
        for q, expand in cases.YEAR_QUESTIONS:
            out["years"].append(ref.extract_years_from_question(q, expand_window=expand))
        b = ref.BM25Retriever(paths["bm25_db_dir"], paths["bm25_documents_dir"], subset_path=paths["subset_path"])
        for name, kw in cases.BM25_CASES.items():
            out["bm25"][name] = cases.normalise(b.retrieve_by_company_name(**kw))
        try:
            b.retrieve_by_company_name("nobody", "q")
            out["errors"]["bm25_unknown_company"] = None
        except Exception as e:  # noqa: BLE001
            out["errors"]["bm25_unknown_company"] = [type(e).__name__, str(e)]
    dst = os.path.join(ROOT, "tests", "golden", "reference_outputs.json")
    with open(dst, "w", encoding="utf-8") as f:
        json.dump(out, f, ensure_ascii=False, sort_keys=True, separators=(",", ":"))
    print("wrote", dst, os.path.getsize(dst), "bytes")


if __name__ == "__main__":
    main()
