"""Caller-level conformance: the REAL caller of the retrieval path — the reference's
`src/questions_processing.py` (QuestionsProcessor, unmodified, imported from /root/reference) — driven
over the miniature corpus of corpus.py with one of two `src.retrieval` modules underneath it:

  --arm reference   the reference's own src/retrieval.py (faiss / rank_bm25 replaced by the numpy oracle,
                    exactly as make_golden.py does) -> writes tests/golden/caller_outputs.json
  --arm ours        this repo's retrievers installed as `src.retrieval` (the re-export INTEGRATION.md §2
                    shows), index objects oracle-backed so the run needs no GPU   (--gpu: the CUDA indexes)

Both arms share the scripted services of stubs.py plus a scripted `APIProcessor` (page selection,
answer generation, question rephrasing).  What is compared is everything QuestionsProcessor builds
from the retriever's output: page groups, page selection context, RAG context, references,
all_retrieved_chunks, initial_retrieval_results, algorithm_contribution, per-question submission records.
Covers SURVEY §8(f) f1 — "callers keep working unchanged against our classes"
(src/questions_processing.py:610-694,734-758,904-986,1042-1138,1334-1399).

/root/reference exists only in the authoring container, so the test that runs the "ours" arm
(tests/test_caller_conformance.py) skips when it is absent; the golden outputs are committed.

    python -m tests.golden.caller_conformance --arm reference          (regenerate the golden file)
    python -m tests.golden.caller_conformance --arm ours --out /tmp/x.json
"""
import argparse
import hashlib
import io
import json
import os
import re
import sys
import tempfile
import types
from contextlib import redirect_stdout

REF = "/root/reference"
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from tests.golden import corpus, stubs  # noqa: E402

GOLDEN = os.path.join(ROOT, "tests", "golden", "caller_outputs.json")
MQ_TWO = {"synonym": True, "subquestion": True, "variant": False}


# --------------------------------------------------------------------------------------------------
# scripted APIProcessor (src/api_requests.py: APIProcessor surface used by QuestionsProcessor)
# --------------------------------------------------------------------------------------------------
def _pages_of(context):
    return [int(p) for p in re.findall(r"Text retrieved from page (\d+):", context)]


class _Selector:
    """`openai_processor.processor.send_message(... response_format=PageSelectionSchema)`
    (src/questions_processing.py:796-802): picks every second retrieved page, plus one page that was
    never retrieved (the caller must drop it, :812-813)."""

    def send_message(self, model=None, system_content="", human_content="", is_structured=False, response_format=None, **kw):
        pages = _pages_of(human_content)
        return {"selected_pages": pages[::2][:5] + [9999], "reasoning": f"scripted: {len(pages)} pages offered"}


class ScriptedAPIProcessor:
    def __init__(self, provider="qwen"):
        self.provider = provider
        self.processor = _Selector()
        self.response_data = {"model": "scripted", "input_tokens": 0, "output_tokens": 0}

    def _build_rag_context_prompts(self, schema):
        return f"system prompt for {schema}", None, "Context:\n{context}\n\nQuestion: {question}"

    def get_answer_from_rag_context(self, question, rag_context, schema, model):
        if schema == "comparative":
            parts = {c: a.get("final_answer") for c, a in sorted(rag_context.items())}
            return {"step_by_step_analysis": "scripted", "reasoning_summary": "comparative", "relevant_pages": [],
                    "final_answer": json.dumps(parts, ensure_ascii=False, sort_keys=True)}
        pages = _pages_of(rag_context)
        digest = hashlib.sha1((question + rag_context).encode("utf-8")).hexdigest()[:16]
        return {"step_by_step_analysis": "scripted", "reasoning_summary": f"{len(pages)} pages in context",
                "relevant_pages": pages[:4], "final_answer": digest}

    def get_rephrased_questions(self, original_question, companies):
        return {c: f"{c} order backlog and export revenue" for c in companies}


# --------------------------------------------------------------------------------------------------
# module stand-ins
# --------------------------------------------------------------------------------------------------
def _install_services():
    api = types.ModuleType("src.api_requests")
    api.BaseQwenProcessor = stubs.ScriptedQwen
    api.BaseGeminiProcessor = object
    api.APIProcessor = ScriptedAPIProcessor
    sys.modules["src.api_requests"] = api
    rer = types.ModuleType("src.reranking")
    rer.LLMReranker = stubs.ScriptedReranker
    sys.modules["src.reranking"] = rer


def install_reference_arm():
    """The reference's src.retrieval over oracle-backed faiss / rank_bm25 stand-ins (as make_golden.py)."""
    from tests.golden import make_golden

    make_golden.install_stubs()
    _install_services()
    sys.path.insert(0, REF)
    import src.retrieval as mod

    assert mod.__file__.startswith(REF), mod.__file__


def install_ours_arm(gpu=False):
    """This repo's retrievers as `src.retrieval` — the re-export of INTEGRATION.md §2."""
    _install_services()
    sys.path.insert(0, REF)
    import rag_challenge_2_b200.retrieval as R

    if not gpu:
        from tests import golden_util as gu

        R.flat_index.read_index = lambda path, device=0: stubs.read_ixfi(path)
        R.load_bm25_pickle = gu.oracle_bm25_loader
    R.clear_resident_cache()
    shim = types.ModuleType("src.retrieval")
    for name in ("VectorRetriever", "BM25Retriever", "HybridRetriever", "route_reports_by_time",
                 "extract_years_from_question"):
        setattr(shim, name, getattr(R, name))
    sys.modules["src.retrieval"] = shim


# --------------------------------------------------------------------------------------------------
# scenarios
# --------------------------------------------------------------------------------------------------
def _processor(QP, paths, **kw):
    base = dict(vector_db_dir=paths["vector_db_dir"], documents_dir=paths["documents_dir"],
                subset_path=paths["subset_path"], new_challenge_pipeline=True, parallel_requests=1,
                use_hyde=False, use_multi_query=False, top_n_retrieval=10)
    base.update(kw)
    return QP(**base)


SCENARIOS = {
    # name -> (QuestionsProcessor kwargs, get_answer_for_company kwargs)
    "vector_parent_pages": (dict(parent_document_retrieval=True),
                            dict(company_name=corpus.COMPANY, question=corpus.QUESTIONS[0], schema="number")),
    "vector_chunks_years_history": (dict(parent_document_retrieval=False, top_n_retrieval=8),
                                    dict(company_name=corpus.COMPANY, question=corpus.QUESTIONS[1], schema="name",
                                         selected_years=[2023, 2024],
                                         conversation_history=[{"question": "earlier question", "answer": "earlier answer"}])),
    "vector_expanded_queries": (dict(parent_document_retrieval=True, use_hyde=True, use_multi_query=True,
                                     multi_query_methods=MQ_TWO, top_n_retrieval=12),
                                dict(company_name=corpus.COMPANY, question=corpus.QUESTIONS[3], schema="names",
                                     selected_years=[2025])),
    "rerank_upstream_groups": (dict(parent_document_retrieval=True, llm_reranking=True, llm_reranking_sample_size=20,
                                    top_n_retrieval=6, expand_upstream=True, expand_top_k=3, expand_context_size=1),
                               dict(company_name=corpus.COMPANY, question=corpus.QUESTIONS[2], schema="boolean",
                                    selected_years=[2024])),
    "rerank_downstream": (dict(parent_document_retrieval=True, llm_reranking=True, llm_reranking_sample_size=15,
                               top_n_retrieval=5),
                          dict(company_name=corpus.OTHER, question=corpus.QUESTIONS[2], schema="number")),
    "hybrid_expansion_method": (dict(parent_document_retrieval=True, retrieval_method="hybrid_expansion", max_hops=2,
                                     neighbor_k=6, top_n_retrieval=10),
                                dict(company_name=corpus.COMPANY, question=corpus.QUESTIONS[0], schema="number",
                                     selected_years=[2024])),
    "triangulation_method_rerank": (dict(parent_document_retrieval=False, llm_reranking=True, llm_reranking_sample_size=12,
                                         retrieval_method="triangulation", max_hops=3, neighbor_k=7, top_n_retrieval=6),
                                    dict(company_name=corpus.COMPANY, question=corpus.QUESTIONS[1], schema="name",
                                         selected_years=[2022, 2023])),
    "full_context": (dict(full_context=True),
                     dict(company_name=corpus.OTHER, question=corpus.QUESTIONS[2], schema="number")),
}

BATCH_QUESTIONS = [
    {"text": f"{corpus.COMPANY} cash flow dividend 2024", "kind": "number"},
    {"question": f"What is the tax profit of {corpus.OTHER} in 2021", "schema": "number"},
    {"text": f"Compare order backlog of {corpus.OTHER} and {corpus.COMPANY} please", "kind": "name"},  # comparative
    {"text": "a question that names no company at all", "kind": "number"},                            # error record
    {"text": "   ", "kind": "number"},                                                                # skipped
]


def canon(obj, key=None):
    """JSON-comparable form: timings zeroed, set-derived lists sorted, provider counters and the
    glossary-dependent prompt text (src/financial_glossary.py — prompt text, not retrieval) dropped."""
    import numpy as np

    if key in ("timing", "reranker_stats", "glossary_context", "concepts", "error_traceback"):
        return None
    if isinstance(obj, dict):
        return {str(k): canon(v, str(k)) for k, v in obj.items()}
    if isinstance(obj, (list, tuple)):
        out = [canon(v) for v in obj]
        if key == "retrieval_sources":
            out = sorted(out)
        if key == "references":  # comparative questions collect them in thread-completion order (:1364-1378)
            out = sorted(out, key=lambda r: json.dumps(r, sort_keys=True, ensure_ascii=False))
        return out
    if isinstance(obj, np.floating):
        return float(obj)
    if isinstance(obj, np.integer):
        return int(obj)
    return obj


def run_scenarios(verbose=False):
    import src.questions_processing as qp_mod

    assert qp_mod.__file__.startswith(REF), qp_mod.__file__
    QP = qp_mod.QuestionsProcessor
    out = {"answers": {}, "errors": {}}
    sink = sys.stderr if verbose else io.StringIO()
    with tempfile.TemporaryDirectory() as tmp, redirect_stdout(sink):
        paths = corpus.build(tmp)
        for name, (init_kw, call_kw) in SCENARIOS.items():
            stages = []
            p = _processor(QP, paths, **init_kw)
            ans = p.get_answer_for_company(progress_callback=lambda stage, pct: stages.append([stage, pct]), **call_kw)
            out["answers"][name] = canon(ans)
            out["answers"][name]["_progress"] = stages
        # error behaviour seen by the caller (src/retrieval.py:844 -> ValueError)
        p = _processor(QP, paths)
        try:
            p.get_answer_for_company(company_name="nobody", question="q", schema="number")
            out["errors"]["unknown_company"] = None
        except Exception as e:  # noqa: BLE001
            out["errors"]["unknown_company"] = [type(e).__name__, str(e)]
        # the batch runner (:1104-1138): per-question records, a comparative question, an error record
        p = _processor(QP, paths, parent_document_retrieval=True, top_n_retrieval=6)
        res = p.process_questions_list([dict(q) for q in BATCH_QUESTIONS])
        out["batch"] = canon(res)
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--arm", choices=["reference", "ours"], required=True)
    ap.add_argument("--gpu", action="store_true", help="ours arm: CUDA index objects instead of oracle-backed ones")
    ap.add_argument("--out", default=None)
    ap.add_argument("--verbose", action="store_true")
    args = ap.parse_args()
    if args.arm == "reference":
        install_reference_arm()
    else:
        install_ours_arm(gpu=args.gpu)
    out = run_scenarios(args.verbose)
    dst = args.out or (GOLDEN if args.arm == "reference" else None)
    text = json.dumps(out, ensure_ascii=False, sort_keys=True, separators=(",", ":"))
    if dst:
        with open(dst, "w", encoding="utf-8") as f:
            f.write(text)
        print("wrote", dst, len(text), "bytes", file=sys.stderr)
    else:
        print(text)


if __name__ == "__main__":
    main()
