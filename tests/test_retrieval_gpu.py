"""GPU: the same golden scenarios as tests/test_retrieval_golden.py, but with the real GPU index
objects (libb2r) under this repo's retriever classes.  fp32 summation order differs from the oracle
(~1e-7 relative), and records carry scores rounded to 4 dp, so scores are compared with a 2e-4
absolute band; membership and every other field must match exactly."""
import pytest

from tests import golden_util as gu
from tests.golden import cases, corpus, stubs

pytestmark = pytest.mark.gpu

BAND = 2e-4


@pytest.fixture(scope="module")
def env(tmp_path_factory):
    import rag_challenge_2_b200.retrieval as R

    R.clear_resident_cache()
    paths = corpus.build(tmp_path_factory.mktemp("db"))
    yield R, paths, gu.load_golden()
    R.clear_resident_cache()


def rec_key(rec):
    return (rec["source_sha1"], rec["page"], rec["text"], rec.get("hit_count"),
            tuple(sorted(rec.get("retrieval_sources", []))), tuple(rec.get("query_sources", [])))


def assert_records_close(got, want, score_fields=("vector_similarity", "max_original_similarity", "combined_score"),
                         sort_field="vector_similarity"):
    assert len(got) == len(want)
    gk, wk = sorted(map(rec_key, got)), sorted(map(rec_key, want))
    assert gk == wk, "different records retrieved"
    wmap = {}
    for rec in want:
        wmap.setdefault(rec_key(rec), []).append(rec)
    for rec in got:
        ref = wmap[rec_key(rec)].pop(0)
        for f in score_fields:
            if f in ref:
                assert abs(rec[f] - ref[f]) <= BAND, (f, rec[f], ref[f])
        assert set(rec.keys()) == set(ref.keys())
    vals = [r[sort_field] for r in got]
    assert all(a >= b for a, b in zip(vals, vals[1:])), "not sorted descending"


@pytest.mark.parametrize("name", sorted(cases.VECTOR_CASES))
def test_vector_retriever_gpu(env, name):
    R, paths, golden = env
    got = gu.canon_result(gu.run_vector_case(R, paths, name, stubs.ScriptedQwen()))
    want = gu.canon_result(golden["vector"][name])
    assert_records_close(got["results"], want["results"])
    assert_records_close(got["initial_retrieval_results"], want["initial_retrieval_results"])
    assert got["expansion_texts"] == want["expansion_texts"]
    assert (got["retrieval_details"] is None) == (want["retrieval_details"] is None)
    if want["retrieval_details"]:
        assert got["retrieval_details"]["method"] == want["retrieval_details"]["method"]
        gi, wi = got["retrieval_details"]["traversal_info"], want["retrieval_details"]["traversal_info"]
        gi, wi = (gi if isinstance(gi, list) else [gi]), (wi if isinstance(wi, list) else [wi])
        assert [t["path"] for t in gi] == [t["path"] for t in wi]   # identical walks
    if want["algorithm_contribution"]:
        for f in ("basic_retrieval_count", "ssg_new_chunks_count", "triangulation_new_chunks_count", "ssg_stats",
                  "triangulation_stats"):
            assert got["algorithm_contribution"][f] == want["algorithm_contribution"][f]


@pytest.mark.parametrize("name", sorted(cases.HYBRID_CASES))
def test_hybrid_retriever_gpu(env, name):
    R, paths, golden = env
    got = gu.canon_result(gu.run_hybrid_case(R, paths, name, stubs.ScriptedQwen()))
    want = gu.canon_result(golden["hybrid"][name])
    assert_records_close(got["results"], want["results"], sort_field="combined_score")
    assert_records_close(got["initial_retrieval_results"], want["initial_retrieval_results"])


@pytest.mark.parametrize("name", sorted(cases.BM25_CASES))
def test_bm25_retriever_gpu(env, name):
    R, paths, golden = env
    b = R.BM25Retriever(paths["bm25_db_dir"], paths["bm25_documents_dir"], subset_path=paths["subset_path"])
    got = cases.normalise(b.retrieve_by_company_name(**cases.BM25_CASES[name]))
    assert_records_close(got, golden["bm25"][name])


def test_retrieve_all_and_threads(env):
    """Whole retrievers used concurrently from several threads (src/questions_processing.py:1123-1125)."""
    from concurrent.futures import ThreadPoolExecutor

    R, paths, golden = env
    r = R.VectorRetriever(paths["vector_db_dir"], paths["documents_dir"], subset_path=paths["subset_path"],
                          qwen=stubs.ScriptedQwen())
    assert cases.normalise(r.retrieve_all(corpus.OTHER)) == golden["retrieve_all"]
    names = ["basic_plain", "basic_years_3", "ssg", "triangulation"] * 3

    def run(name):
        return gu.canon_result(gu.run_vector_case(R, paths, name, stubs.ScriptedQwen()))

    with ThreadPoolExecutor(max_workers=6) as ex:
        outs = list(ex.map(run, names))
    for name, got in zip(names, outs):
        assert_records_close(got["results"], gu.canon_result(golden["vector"][name])["results"])


def test_retrieve_batch_on_resident_year_dbs(b2r):
    """Config-2 shape at CI size: six synthetic year-DBs resident in HBM (lazy chunk / page records), a
    question batch routed by year; retrieve_batch (one tensor-core batch per DB) == per-question calls."""
    import torch

    from rag_challenge_2_b200 import retrieval, synth

    years = [2020, 2021, 2022, 2023, 2024, 2025]
    reports = []
    for i, y in enumerate(years):
        lo = i * 4 * synth.ROW_BLOCK
        rows = synth.dense_rows_torch(lo, lo + 20000 + 3000 * i, 1024, torch.device("cuda", 0))
        reports.append(synth.synthetic_year_report(f"J{y}", "金盘科技", rows.shape[0], b2r.FlatIPIndex.from_device_tensor(rows)))
    r = retrieval.VectorRetriever.from_reports(reports, qwen=retrieval.HashEmbedder(1024))
    qs = [f"{years[i % 6]}年的第{i}个指标" for i in range(40)]
    sel = [None if i % 3 == 0 else ([2024] if i % 3 == 1 else [2023, 2024, 2025]) for i in range(40)]
    got = r.retrieve_batch("金盘科技", qs, top_n=30, selected_years=sel, use_hyde=False, use_multi_query=False)
    for q, y, g in zip(qs, sel, got):
        want = r.retrieve_by_company_name("金盘科技", q, top_n=30, selected_years=y, use_hyde=False, use_multi_query=False)
        assert_records_close(g["results"], want["results"])
        n_dbs = 6 if y is None else len(y)
        assert len(g["initial_retrieval_results"]) == 30 * n_dbs
        assert {rec["source_sha1"] for rec in g["results"]} <= {f"J{v}" for v in (y or years)}
