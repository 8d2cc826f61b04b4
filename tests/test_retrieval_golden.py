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
