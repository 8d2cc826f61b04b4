"""Differential run: the REAL reference retrievers (/root/reference/src/retrieval.py, unmodified, over the
oracle-backed faiss / rank_bm25 stand-ins of make_golden.py) and this repo's retrievers (same stand-in
index objects), side by side in one process, on randomly drawn call parameters over the miniature corpus:
top_n, parent pages, year filters, HyDE / multi-query switches and method subsets, retrieval method, hop
and neighbour counts, reranking sample sizes, BM25 queries.  Every returned dictionary must be identical
(timings zeroed, set-derived lists sorted — tests/golden_util.canon_result).

Complements the fixed scenarios of cases.py: the committed golden file pins 25 hand-picked calls, this
sweeps the parameter space (seeded, so a failure reproduces).  Needs the reference checkout, i.e. the
authoring container; tests/test_differential.py skips elsewhere.

    python -m tests.golden.differential --cases 200 --seed 1
"""
import argparse
import io
import json
import os
import sys
import tempfile
from contextlib import redirect_stdout

import numpy as np

REF = "/root/reference"
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from tests import golden_util as gu  # noqa: E402
from tests.golden import corpus, make_golden, stubs  # noqa: E402

WORDS = corpus.TOPICS + corpus.FILL + ["2020", "2021", "2022", "2023", "2024", "2025", "金盘科技", "毛利率", "zzzz"]


def draw_vector_case(rng):
    method = str(rng.choice(["basic", "basic", "ssg", "triangulation", "hybrid_expansion", "nonsense"]))
    company = corpus.COMPANY if rng.random() < 0.85 else corpus.OTHER
    if rng.random() < 0.03:
        company = "nobody"  # ValueError("No report found with ...") on both sides (:844)
    years = None
    r = rng.random()
    if r < 0.5:
        years = sorted({int(y) for y in rng.choice(np.arange(2019, 2027), size=int(rng.integers(1, 4)))})
    elif r < 0.55:
        years = []
    use_mq = bool(rng.random() < 0.35)
    mq = None
    if use_mq and rng.random() < 0.8:
        mq = {m: bool(rng.random() < 0.6) for m in ("synonym", "subquestion", "variant")}
    query = " ".join(str(w) for w in rng.choice(WORDS, size=int(rng.integers(1, 9))))
    kw = dict(company_name=company, query=query, top_n=int(rng.choice([1, 2, 3, 5, 10, 30, 64, 100])),
              return_parent_pages=bool(rng.random() < 0.5), use_hyde=bool(rng.random() < 0.3), use_multi_query=use_mq,
              multi_query_config=mq, selected_years=years, retrieval_method=method,
              max_hops=int(rng.integers(1, 5)), neighbor_k=int(rng.choice([1, 3, 6, 10, 30, 80])))
    if rng.random() < 0.3:
        kw["llm_reranking_sample_size"] = int(rng.integers(1, 40))
    return kw


def draw_hybrid_case(rng):
    kw = draw_vector_case(rng)
    kw["llm_reranking_sample_size"] = int(rng.integers(1, 40))
    kw["documents_batch_size"] = int(rng.integers(1, 6))
    kw["llm_weight"] = float(rng.choice([0.3, 0.7]))
    return kw


def draw_bm25_case(rng):
    query = " ".join(str(w) for w in rng.choice(WORDS, size=int(rng.integers(1, 10))))
    years = None
    if rng.random() < 0.5:
        years = sorted({int(y) for y in rng.choice(np.arange(2019, 2027), size=int(rng.integers(1, 4)))})
    return dict(company_name=corpus.COMPANY if rng.random() < 0.8 else corpus.OTHER, query=query,
                top_n=int(rng.choice([1, 3, 7, 30, 200])), return_parent_pages=bool(rng.random() < 0.5),
                selected_years=years)


def corrupt(paths):
    """Damage the fixture the way real data directories get damaged: a page record missing (hits on its chunks fail,
    the hits before them are kept, :1145,1441), an index with more rows than the report has chunks (ids past the chunk
    list), and one with fewer (faiss pads with id -1, which Python indexes as the LAST chunk — kept as is)."""
    import json

    docs, vec = paths["documents_dir"], paths["vector_db_dir"]
    with open(docs / "J2023.json", encoding="utf-8") as f:
        rep = json.load(f)
    rep["content"]["pages"] = [p for p in rep["content"]["pages"] if p["page"] not in (3, 7)]
    with open(docs / "J2023.json", "w", encoding="utf-8") as f:
        json.dump(rep, f, ensure_ascii=False)
    rows = stubs.read_ixfi(vec / "J2024.faiss").rows
    corpus.write_ixfi(vec / "J2024.faiss", np.concatenate([rows, rows[:7] * np.float32(1.01)]))
    rows = stubs.read_ixfi(vec / "J2025.faiss").rows
    corpus.write_ixfi(vec / "J2025.faiss", rows[:9])


def _first_differences(x, y, path="", out=None, limit=6):
    out = [] if out is None else out
    if len(out) >= limit:
        return out
    if type(x) is not type(y):
        out.append(f"{path}: {type(x).__name__} vs {type(y).__name__}")
    elif isinstance(x, dict):
        for k in sorted(set(x) | set(y)):
            if k not in x or k not in y:
                out.append(f"{path}/{k}: only on one side")
            else:
                _first_differences(x[k], y[k], f"{path}/{k}", out, limit)
    elif isinstance(x, list):
        if len(x) != len(y):
            out.append(f"{path}: length {len(x)} vs {len(y)}")
        else:
            for j, (p, q) in enumerate(zip(x, y)):
                _first_differences(p, q, f"{path}[{j}]", out, limit)
    elif x != y:
        out.append(f"{path}: {x!r:.80} vs {y!r:.80}")
    return out


def _call(fn):
    try:
        return {"ok": fn()}
    except Exception as e:  # noqa: BLE001 - error behaviour is part of the contract
        return {"error": [type(e).__name__, str(e)]}


def _canon(x):
    if "ok" in x:
        return {"ok": gu.canon_result(x["ok"]) if isinstance(x["ok"], dict) else gu.canon_records(gu.strip_unpinned(x["ok"])),
                "progress": x.get("progress")}
    return x


def run(n_cases, seed, verbose=False, corrupted=False):
    make_golden.install_stubs()
    sys.path.insert(0, REF)
    import src.retrieval as ref

    assert ref.__file__.startswith(REF), ref.__file__
    import rag_challenge_2_b200.retrieval as ours

    ours.flat_index.read_index = lambda path, device=0: stubs.read_ixfi(path)
    ours.load_bm25_pickle = gu.oracle_bm25_loader
    ours.clear_resident_cache()
    rng = np.random.default_rng(seed)
    failures = []
    sink = sys.stderr if verbose else io.StringIO()
    with tempfile.TemporaryDirectory() as tmp, redirect_stdout(sink):
        paths = corpus.build(tmp)
        if corrupted:
            corrupt(paths)
        common = dict(vector_db_dir=paths["vector_db_dir"], documents_dir=paths["documents_dir"],
                      subset_path=paths["subset_path"], parallel_workers=1)
        base_common = dict(common)
        attempts = {}
        i = -1
        while i + 1 < n_cases:
            i += 1
            # The reference collects its per-(report, query) tasks with as_completed (:1506), which yields futures that
            # are already finished in set order: even with one worker the aggregation order — and with it the order of
            # equal scores and of a record's query_sources — is occasionally different from the submission order.  A
            # mismatch is therefore replayed (same parameters: the generator state is restored) and counts only if it
            # persists three times; a real divergence always does.
            state = rng.bit_generator.state
            kind = ("vector", "vector", "vector", "hybrid", "bm25")[i % 5]
            # year routing with and without subset.csv (metainfo.year, then the sha1_name pattern, then the csv: :101-129)
            common = dict(base_common, subset_path=(paths["subset_path"] if rng.random() < 0.75 else None))
            if i % 10 == 7:
                # this repo's retrieve_batch (one batched search per report for the whole question list) against the
                # reference called once per question (its caller's loop, src/questions_processing.py:1104-1138)
                kind = "batch"
                kw = draw_vector_case(rng)
                kw.pop("llm_reranking_sample_size", None)
                kw.pop("query")
                if kw["company_name"] == "nobody":
                    kw["company_name"] = corpus.COMPANY
                questions = [" ".join(str(w) for w in rng.choice(WORDS, size=int(rng.integers(1, 9))))
                             for _ in range(int(rng.integers(2, 7)))]
                a = _call(lambda: [ref.VectorRetriever(**common).retrieve_by_company_name(query=q, **kw) for q in questions])
                b = _call(lambda: ours.VectorRetriever(**common, qwen=stubs.ScriptedQwen()).retrieve_batch(
                    kw["company_name"], questions, **{k: v for k, v in kw.items() if k != "company_name"}))
                for side in (a, b):
                    if "ok" in side:
                        side["ok"] = {"results": [gu.canon_result(r) for r in side["ok"]]}
                        for r in side["ok"]["results"]:
                            r["timing"] = None
            elif i % 20 == 3:
                # HybridRetriever.retrieve_batch against the reference called per question
                kind = "hybrid_batch"
                kw = draw_hybrid_case(rng)
                kw.pop("query")
                if kw["company_name"] == "nobody":
                    kw["company_name"] = corpus.COMPANY
                questions = [" ".join(str(w) for w in rng.choice(WORDS, size=int(rng.integers(1, 9))))
                             for _ in range(int(rng.integers(2, 5)))]
                a = _call(lambda: [ref.HybridRetriever(**common).retrieve_by_company_name(query=q, **kw) for q in questions])
                b = _call(lambda: ours.HybridRetriever(**common, qwen=stubs.ScriptedQwen(), reranker=stubs.ScriptedReranker())
                          .retrieve_batch(kw["company_name"], questions, **{k: v for k, v in kw.items() if k != "company_name"}))
                for side in (a, b):
                    if "ok" in side:
                        side["ok"] = {"per_question": [gu.canon_result(r) for r in side["ok"]]}
                        for r in side["ok"]["per_question"]:
                            r["timing"] = None
            elif kind == "vector":
                kw = draw_vector_case(rng)
                init = {}
                if rng.random() < 0.2:  # constructor defaults that the call's defaults override or not (:808-822)
                    init = dict(retrieval_method=kw.pop("retrieval_method"), max_hops=kw.pop("max_hops"),
                                neighbor_k=kw.pop("neighbor_k"))
                log_a, log_b = [], []   # the progress callback's (stage, percent) sequence is part of the comparison
                a = _call(lambda: ref.VectorRetriever(**common, **init).retrieve_by_company_name(
                    progress_callback=lambda stage, pct: log_a.append([stage, pct]), **kw))
                b = _call(lambda: ours.VectorRetriever(**common, qwen=stubs.ScriptedQwen(), **init).retrieve_by_company_name(
                    progress_callback=lambda stage, pct: log_b.append([stage, pct]), **kw))
                a["progress"], b["progress"] = log_a, log_b
            elif kind == "hybrid":
                kw = draw_hybrid_case(rng)
                log_a, log_b = [], []
                a = _call(lambda: ref.HybridRetriever(**common).retrieve_by_company_name(
                    progress_callback=lambda stage, pct: log_a.append([stage, pct]), **kw))
                b = _call(lambda: ours.HybridRetriever(**common, qwen=stubs.ScriptedQwen(),
                                                       reranker=stubs.ScriptedReranker()).retrieve_by_company_name(
                    progress_callback=lambda stage, pct: log_b.append([stage, pct]), **kw))
                a["progress"], b["progress"] = log_a, log_b
            elif i % 10 == 9:
                # BM25Retriever.retrieve_batch (one fused launch per report) against the reference called per question,
                # with per-question year filters
                kind = "bm25_batch"
                kw = draw_bm25_case(rng)
                n_q = int(rng.integers(2, 7))
                questions = [" ".join(str(w) for w in rng.choice(WORDS, size=int(rng.integers(1, 10)))) for _ in range(n_q)]
                years = [draw_bm25_case(rng)["selected_years"] for _ in range(n_q)]
                args = (paths["bm25_db_dir"], paths["bm25_documents_dir"])
                a = _call(lambda: {"per_question": [ref.BM25Retriever(*args, subset_path=paths["subset_path"]).retrieve_by_company_name(
                    kw["company_name"], q, top_n=kw["top_n"], return_parent_pages=kw["return_parent_pages"], selected_years=y)
                    for q, y in zip(questions, years)]})
                b = _call(lambda: {"per_question": ours.BM25Retriever(*args, subset_path=paths["subset_path"]).retrieve_batch(
                    kw["company_name"], questions, top_n=kw["top_n"], return_parent_pages=kw["return_parent_pages"],
                    selected_years=years)})
            else:
                kw = draw_bm25_case(rng)
                args = (paths["bm25_db_dir"], paths["bm25_documents_dir"])
                a = _call(lambda: ref.BM25Retriever(*args, subset_path=paths["subset_path"]).retrieve_by_company_name(**kw))
                b = _call(lambda: ours.BM25Retriever(*args, subset_path=paths["subset_path"]).retrieve_by_company_name(**kw))
            ca, cb = _canon(a), _canon(b)
            ja = json.dumps(ca, sort_keys=True, ensure_ascii=False)
            jb = json.dumps(cb, sort_keys=True, ensure_ascii=False)
            if ja != jb and attempts.get(i, 0) < 2:
                attempts[i] = attempts.get(i, 0) + 1
                rng.bit_generator.state = state
                i -= 1
                continue
            if ja != jb:
                failures.append({"case": i, "kind": kind, "kwargs": kw, "attempts": attempts.get(i, 0) + 1,
                                 "questions": questions if kind.endswith("batch") else None,
                                 "first_differences": _first_differences(ca, cb), "reference": ja[:400], "ours": jb[:400]})
    run.replayed = sorted(attempts)
    return failures


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--cases", type=int, default=100)
    ap.add_argument("--seed", type=int, default=1)
    ap.add_argument("--verbose", action="store_true")
    ap.add_argument("--corrupt", action="store_true", help="damaged fixture: missing pages, index / chunk count mismatches")
    args = ap.parse_args()
    failures = run(args.cases, args.seed, args.verbose, args.corrupt)
    print(json.dumps({"cases": args.cases, "seed": args.seed, "replayed_cases": run.replayed, "failures": failures},
                     ensure_ascii=False, indent=1))
    raise SystemExit(1 if failures else 0)


if __name__ == "__main__":
    main()
