"""Helpers shared by the golden-fixture tests (CPU exact / GPU tolerant)."""
import json
import os
import pickle

from tests.golden import cases, corpus, stubs

GOLDEN_PATH = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "reference_outputs.json")


def load_golden():
    with open(GOLDEN_PATH, "r", encoding="utf-8") as f:
        return json.load(f)


class _OracleBM25Unpickler(pickle.Unpickler):
    def find_class(self, module, name):
        if module == "rank_bm25":
            return stubs.OracleBM25
        return super().find_class(module, name)


def oracle_bm25_loader(path, device=0):
    with open(str(path), "rb") as f:
        return _OracleBM25Unpickler(f).load()


def strip_unpinned(result):
    """Drop what depends on the reference's glossary module (prompt text, not retrieval)."""
    r = json.loads(json.dumps(cases.normalise(result), ensure_ascii=False))
    et = r.get("expansion_texts") if isinstance(r, dict) else None
    if et:
        et["glossary_context"] = None
        for t in et.get("multi_query_texts", []):
            t["concepts"] = None
    return r


def canon_records(records):
    out = []
    for rec in records:
        rec = dict(rec)
        if "retrieval_sources" in rec:
            rec["retrieval_sources"] = sorted(rec["retrieval_sources"])
        out.append(rec)
    return out


def canon_result(r):
    r = strip_unpinned(r)
    if isinstance(r, dict):
        for key in ("results", "initial_retrieval_results"):
            if key in r:
                r[key] = canon_records(r[key])
        if r.get("reranker_stats") is not None:
            r["reranker_stats"] = None  # provider-specific counters
    return r


def run_vector_case(retrieval_mod, paths, name, qwen):
    kw = dict(cases.VECTOR_CASES[name])
    init = kw.pop("_init", {})
    r = retrieval_mod.VectorRetriever(paths["vector_db_dir"], paths["documents_dir"], subset_path=paths["subset_path"],
                                      parallel_workers=1, qwen=qwen, **init)
    return r.retrieve_by_company_name(**kw)


def run_hybrid_case(retrieval_mod, paths, name, qwen):
    r = retrieval_mod.HybridRetriever(paths["vector_db_dir"], paths["documents_dir"], subset_path=paths["subset_path"],
                                      parallel_workers=1, qwen=qwen, reranker=stubs.ScriptedReranker())
    return r.retrieve_by_company_name(**cases.HYBRID_CASES[name])
