"""CPU: this repo's retriever classes (host logic: routing, expansion protocol, hit records,
aggregation, score fusion, sort, cut, BM25 glue) against golden outputs produced by the REAL
reference code (tests/golden/make_golden.py).  The index objects are oracle-backed here, so the
records must match EXACTLY.  The GPU-backed comparison is tests/test_retrieval_gpu.py."""
import json

import pytest

from tests import golden_util as gu
from tests.golden import cases, corpus, stubs


@pytest.fixture(scope="module")
def env(tmp_path_factory):
    import rag_challenge_2_b200.retrieval as R

    paths = corpus.build(tmp_path_factory.mktemp("db"))
    mp = pytest.MonkeyPatch()
    mp.setattr(R.flat_index, "read_index", lambda path, device=0: stubs.read_ixfi(path))
    mp.setattr(R, "load_bm25_pickle", gu.oracle_bm25_loader)
    R.clear_resident_cache()
    yield R, paths, gu.load_golden()
    mp.undo()
    R.clear_resident_cache()


@pytest.mark.parametrize("name", sorted(cases.VECTOR_CASES))
def test_vector_retriever_matches_reference(env, name):
    R, paths, golden = env
    got = gu.canon_result(gu.run_vector_case(R, paths, name, stubs.ScriptedQwen()))
    want = gu.canon_result(golden["vector"][name])
    assert list(got.keys()) == ["results", "timing", "expansion_texts", "retrieval_details",
                                "initial_retrieval_results", "algorithm_contribution"]
    assert json.dumps(got, sort_keys=True, ensure_ascii=False) == json.dumps(want, sort_keys=True, ensure_ascii=False)


@pytest.mark.parametrize("name", sorted(cases.HYBRID_CASES))
def test_hybrid_retriever_matches_reference(env, name):
    R, paths, golden = env
    got = gu.canon_result(gu.run_hybrid_case(R, paths, name, stubs.ScriptedQwen()))
    want = gu.canon_result(golden["hybrid"][name])
    assert list(got.keys()) == ["results", "initial_retrieval_results", "timing", "expansion_texts", "reranker_stats",
                                "algorithm_contribution"]
    assert got == want


@pytest.mark.parametrize("name", sorted(cases.BM25_CASES))
def test_bm25_retriever_matches_reference(env, name):
    R, paths, golden = env
    b = R.BM25Retriever(paths["bm25_db_dir"], paths["bm25_documents_dir"], subset_path=paths["subset_path"])
    got = cases.normalise(b.retrieve_by_company_name(**cases.BM25_CASES[name]))
    assert got == golden["bm25"][name]


def test_loading_routing_years_errors(env):
    R, paths, golden = env
    r = R.VectorRetriever(paths["vector_db_dir"], paths["documents_dir"], subset_path=paths["subset_path"],
                          qwen=stubs.ScriptedQwen())
    assert sorted(db["name"] for db in r.all_dbs) == golden["loaded_reports"]
    assert r.year_lookup == golden["year_lookup"]
    assert cases.normalise(r.retrieve_all(corpus.OTHER)) == golden["retrieve_all"]
    for kw, want in zip(cases.ROUTE_CASES, golden["route"]):
        got = R.route_reports_by_time(all_reports=r.all_dbs, **kw)
        assert sorted(rep["name"] for rep in got) == want, kw
    for (q, expand), want in zip(cases.YEAR_QUESTIONS, golden["years"]):
        assert R.extract_years_from_question(q, expand_window=expand) == want
    with pytest.raises(ValueError) as e:
        r.retrieve_by_company_name("nobody", "q", use_hyde=False, use_multi_query=False)
    assert [type(e.value).__name__, str(e.value)] == golden["errors"]["vector_unknown_company"]
    with pytest.raises(ValueError) as e:
        r.retrieve_all("nobody")
    assert str(e.value) == golden["errors"]["retrieve_all_unknown_company"][1]
    b = R.BM25Retriever(paths["bm25_db_dir"], paths["bm25_documents_dir"])
    with pytest.raises(ValueError) as e:
        b.retrieve_by_company_name("nobody", "q")
    assert str(e.value) == golden["errors"]["bm25_unknown_company"][1]


def test_documented_worked_examples():
    """The only known answers the reference itself states: year windows (src/retrieval.py:39-42) and
    the bonus multipliers (docs/retrieval_algorithms/retrieval_algorithms_update_20251203.md:139-151)."""
    import rag_challenge_2_b200.retrieval as R

    assert R.extract_years_from_question("2024年xxx") == [2023, 2024, 2025]
    assert R.extract_years_from_question("2024年相比2023年") == [2022, 2023, 2024, 2025]
    assert R.extract_years_from_question("2024年xxx", expand_window=False) == [2024]
    fuse = R.VectorRetriever.fuse_similarity
    base = {"similarities": [0.5, 0.4], "retrieval_sources": ["basic"]}
    assert fuse(dict(base, count=1))[0] == 0.5
    assert abs(fuse(dict(base, count=2))[0] - 0.5 * 1.2) < 1e-12      # 2 hits -> x1.2
    assert abs(fuse(dict(base, count=3))[0] - 0.5 * 1.4) < 1e-12      # 3 hits -> x1.4
    two = dict(base, count=1, retrieval_sources=["basic", "ssg"])
    three = dict(base, count=1, retrieval_sources=["basic", "ssg", "triangulation"])
    assert abs(fuse(two)[0] - 0.5 * 1.1) < 1e-12                      # 2 methods -> x1.1
    assert abs(fuse(three)[0] - 0.5 * 1.2) < 1e-12                    # 3 methods -> x1.2


def test_embedding_failure_and_progress_callback(env):
    R, paths, _ = env

    class Broken(stubs.ScriptedQwen):
        def get_embeddings(self, texts, model=None):
            raise RuntimeError("offline")

    r = R.VectorRetriever(paths["vector_db_dir"], paths["documents_dir"], qwen=Broken())
    with pytest.raises(ValueError, match="Failed to generate embeddings for all queries."):
        r.retrieve_by_company_name(corpus.COMPANY, "q", use_hyde=False, use_multi_query=False)
    seen = []
    r = R.VectorRetriever(paths["vector_db_dir"], paths["documents_dir"], qwen=stubs.ScriptedQwen())
    r.retrieve_by_company_name(corpus.COMPANY, corpus.QUESTIONS[0], use_hyde=True, use_multi_query=True,
                               progress_callback=lambda stage, pct: seen.append(pct))
    assert seen == [28, 32, 38, 45, 55]      # percentages of src/retrieval.py:832,894,948,1117,1759


def test_resident_cache_shares_reports(env):
    R, paths, _ = env
    a = R.VectorRetriever(paths["vector_db_dir"], paths["documents_dir"], qwen=stubs.ScriptedQwen())
    b = R.VectorRetriever(paths["vector_db_dir"], paths["documents_dir"], qwen=stubs.ScriptedQwen())
    ia = {db["name"]: db["vector_db"] for db in a.all_dbs}
    ib = {db["name"]: db["vector_db"] for db in b.all_dbs}
    assert all(ia[k] is ib[k] for k in ia)   # one resident index per report per process


def _assert_same_retrieval(got, want):
    """Equal up to the oracle index's own batch-size noise: numpy's fp32 GEMV (one query) and GEMM (a batch)
    differ in the last bit, which can flip a score's 4th decimal; everything else must be identical."""
    got, want = gu.canon_result(got), gu.canon_result(want)
    assert got.keys() == want.keys()
    for key in got:
        if key not in ("results", "initial_retrieval_results"):
            assert got[key] == want[key], key
            continue
        assert len(got[key]) == len(want[key])

        def ident(rec):
            return json.dumps({k: v for k, v in rec.items() if k not in ("vector_similarity", "max_original_similarity")},
                              sort_keys=True, ensure_ascii=False)

        g, w = sorted(got[key], key=ident), sorted(want[key], key=ident)
        for a, b in zip(g, w):
            assert ident(a) == ident(b)
            for f in ("vector_similarity", "max_original_similarity"):
                assert abs(a[f] - b[f]) <= 1.5e-4 * max(1.0, a["hit_count"]), (f, a[f], b[f])
        vals = [rec["vector_similarity"] for rec in got[key]]
        assert all(x >= y for x, y in zip(vals, vals[1:]))


def test_retrieve_batch_equals_per_question_calls(env):
    """retrieve_batch (one batched search per report for the whole question batch) must return exactly
    what per-question retrieve_by_company_name calls return — same records, order and scores."""
    R, paths, golden = env
    qs = list(corpus.QUESTIONS) * 2
    years = [None, [2024], [2023, 2024, 2025], [1999], [2022], None, [2020, 2021], [2025]][: len(qs)]
    years += [None] * (len(qs) - len(years))
    r = R.VectorRetriever(paths["vector_db_dir"], paths["documents_dir"], subset_path=paths["subset_path"],
                          qwen=stubs.ScriptedQwen())
    for kw in (dict(use_hyde=False, use_multi_query=False),
               dict(use_hyde=True, use_multi_query=True, multi_query_config=cases.MQ_ALL, return_parent_pages=True)):
        got = r.retrieve_batch(corpus.COMPANY, qs, top_n=10, selected_years=years, **kw)
        assert len(got) == len(qs)
        for q, y, g in zip(qs, years, got):
            want = r.retrieve_by_company_name(corpus.COMPANY, q, top_n=10, selected_years=y, **kw)
            _assert_same_retrieval(g, want)
    # one year list for the whole batch, and a graph method (falls back to the per-question loop)
    got = r.retrieve_batch(corpus.COMPANY, qs[:3], top_n=5, selected_years=[2023, 2024], use_hyde=False,
                           use_multi_query=False, retrieval_method="ssg", max_hops=3, neighbor_k=8)
    for q, g in zip(qs[:3], got):
        want = r.retrieve_by_company_name(corpus.COMPANY, q, top_n=5, selected_years=[2023, 2024], use_hyde=False,
                                          use_multi_query=False, retrieval_method="ssg", max_hops=3, neighbor_k=8)
        assert gu.canon_result(g) == gu.canon_result(want)
    with pytest.raises(ValueError, match="No report found"):
        r.retrieve_batch("nobody", qs[:2], top_n=3, use_hyde=False, use_multi_query=False)


def test_device_placement_round_robin(env):
    """device=[...]: report i (sorted by stem) is loaded on devices[i % n] — one year-DB per GPU."""
    R, paths, golden = env
    seen = {}
    orig = R.flat_index.read_index

    def spy(path, device=0):
        seen[path.stem if hasattr(path, "stem") else str(path)] = device
        return orig(path, device=device)

    mp = pytest.MonkeyPatch()
    mp.setattr(R.flat_index, "read_index", spy)
    R.clear_resident_cache()
    try:
        r = R.VectorRetriever(paths["vector_db_dir"], paths["documents_dir"], subset_path=paths["subset_path"],
                              qwen=stubs.ScriptedQwen(), device=[0, 1, 2])
        stems = sorted(seen)
        assert [seen[s] for s in stems] == [i % 3 for i in range(len(stems))]
        assert r.devices == [0, 1, 2] and r.device == 0
    finally:
        mp.undo()
        R.clear_resident_cache()
    with pytest.raises(ValueError):
        R.VectorRetriever(paths["vector_db_dir"], paths["documents_dir"], device=[])


def test_graph_expansions_use_batched_calls(env):
    """hybrid_expansion walks 30 anchors per (query, report): in lockstep that is 1 + max_hops batched searches
    and a few batched gathers per report, and no per-candidate reconstruct round trips — with the reference's
    records (the golden comparison above runs through the same path)."""
    R, paths, golden = env
    r = R.VectorRetriever(paths["vector_db_dir"], paths["documents_dir"], subset_path=paths["subset_path"],
                          qwen=stubs.ScriptedQwen())
    calls = {"search": [], "reconstruct": 0, "reconstruct_batch": 0}
    cls = type(r.all_dbs[0]["vector_db"])
    orig = cls.search, cls.reconstruct, cls.reconstruct_batch
    mp = pytest.MonkeyPatch()
    mp.setattr(cls, "search", lambda self, x, k: (calls["search"].append(len(x)), orig[0](self, x, k))[1])
    mp.setattr(cls, "reconstruct", lambda self, i: (calls.__setitem__("reconstruct", calls["reconstruct"] + 1), orig[1](self, i))[1])
    mp.setattr(cls, "reconstruct_batch",
               lambda self, ids: (calls.__setitem__("reconstruct_batch", calls["reconstruct_batch"] + 1), orig[2](self, ids))[1])
    try:
        kw = dict(cases.VECTOR_CASES["hybrid_expansion"])
        got = r.retrieve_by_company_name(**kw)
    finally:
        mp.undo()
    want = gu.canon_result(golden["vector"]["hybrid_expansion"])
    assert gu.canon_result(got)["results"] == want["results"]
    n_reports = 1                                   # selected_years=[2024]
    assert calls["reconstruct"] == 0
    assert len(calls["search"]) <= n_reports * (1 + kw["max_hops"])
    assert max(calls["search"]) >= 20               # the walks advance together
    assert calls["reconstruct_batch"] <= n_reports * (2 + kw["max_hops"])


def test_damaged_data_directory_behaves_like_the_reference(tmp_path):
    """Portable restatement of what tests/golden/differential.py --corrupt establishes against the real reference:
    a report whose JSON lost page records keeps the hits ranked before the first failing one (:1441-1442), and a
    hybrid_expansion task that fails in its basic stage ends the call in the reference's UnboundLocalError
    (:1289,1461)."""
    import rag_challenge_2_b200.retrieval as R
    from tests.golden import differential

    paths = corpus.build(tmp_path)
    differential.corrupt(paths)
    mp = pytest.MonkeyPatch()
    mp.setattr(R.flat_index, "read_index", lambda path, device=0: stubs.read_ixfi(path))
    R.clear_resident_cache()
    try:
        r = R.VectorRetriever(paths["vector_db_dir"], paths["documents_dir"], subset_path=paths["subset_path"],
                              parallel_workers=1, qwen=stubs.ScriptedQwen())
        kw = dict(company_name=corpus.COMPANY, query=corpus.QUESTIONS[0], use_hyde=False, use_multi_query=False,
                  selected_years=[2023])
        full = len(r.all_dbs[[d["name"] for d in r.all_dbs].index("J2023")]["document"]["content"]["chunks"])
        partial = 0
        for question in corpus.QUESTIONS:
            got = r.retrieve_by_company_name(top_n=full, **dict(kw, query=question))
            assert len(got["results"]) < full                     # ranks after the first chunk of a lost page are gone
            assert all(rec["page"] not in (3, 7) for rec in got["results"])
            partial += bool(got["results"])
        assert partial                                            # ... and the ranks before it are kept
        with pytest.raises(UnboundLocalError, match="ssg_total_expanded"):
            r.retrieve_by_company_name(top_n=10, retrieval_method="hybrid_expansion", **kw)
        # more index rows than chunks (J2024): ids past the chunk list fail the hit; fewer (J2025): faiss pads with -1,
        # which indexes the last chunk
        r.retrieve_by_company_name(company_name=corpus.COMPANY, query=corpus.QUESTIONS[1], top_n=100, use_hyde=False,
                                   use_multi_query=False, selected_years=[2024, 2025])
    finally:
        mp.undo()
        R.clear_resident_cache()


def test_bm25_retriever_batch_and_document_cache(env, tmp_path):
    """retrieve_batch == per-question calls (per-question year filters); parsed report JSONs are cached per file
    version: an edited report is re-read (the reference re-reads every file on every call, :199-206)."""
    import json as _json
    import os
    import shutil

    R, paths, _ = env
    docs = tmp_path / "docs"
    shutil.copytree(paths["bm25_documents_dir"], docs)
    b = R.BM25Retriever(paths["bm25_db_dir"], docs, subset_path=paths["subset_path"])
    questions = ["gross margin revenue 2024", "cash flow dividend the company", "zzzz", ""]
    years = [None, [2023, 2025], [2024], None]
    got = b.retrieve_batch(corpus.COMPANY, questions, top_n=6, return_parent_pages=True, selected_years=years)
    want = [b.retrieve_by_company_name(corpus.COMPANY, q, top_n=6, return_parent_pages=True, selected_years=y)
            for q, y in zip(questions, years)]
    assert got == want and len(got) == 4 and got[0]
    first = b.retrieve_by_company_name(corpus.COMPANY, questions[0], top_n=3, selected_years=[2024])
    path = docs / "J2024.json"
    with open(path, encoding="utf-8") as f:
        rep = _json.load(f)
    for c in rep["content"]["chunks"]:
        c["text"] = "EDITED " + c["text"]
    with open(path, "w", encoding="utf-8") as f:
        _json.dump(rep, f, ensure_ascii=False)
    os.utime(path, ns=(path.stat().st_atime_ns, path.stat().st_mtime_ns + 1_000_000_000))
    again = b.retrieve_by_company_name(corpus.COMPANY, questions[0], top_n=3, selected_years=[2024])
    assert [r["text"] for r in again] == ["EDITED " + r["text"] for r in first]
    with pytest.raises(ValueError, match="No report found with 'nobody' company name."):
        b.retrieve_batch("nobody", ["q"])
