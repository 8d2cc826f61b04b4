"""N x N cosine analysis against golden statistics produced by the REAL
semantic_similarity_analysis/generate_similarity_matrix.py (tests/golden/make_cosine_golden.py).
CPU: pins oracle/cosine.py (same numpy operations -> equal to float32 rounding).
GPU: this repo's SimilarityMatrixGenerator mirror (K5, no N x N matrix) within the K5 tolerances
(mean / std / min / max 1e-5 relative; median: exact path for n <= 1024, 2e-6 — SURVEY 8d)."""
import json
from pathlib import Path

import pytest

import oracle
from tests.golden import corpus, make_cosine_golden as mk, stubs

STAT_KEYS = ("avg_similarity", "median_similarity", "std_similarity", "max_similarity", "min_similarity")


@pytest.fixture(scope="module")
def env(tmp_path_factory):
    tmp = tmp_path_factory.mktemp("cos")
    paths = corpus.build(tmp)
    scaled = mk.build_scaled_copy(paths["vector_db_dir"], Path(tmp) / "vector_dbs_scaled")
    with open(mk.GOLDEN, encoding="utf-8") as f:
        golden = json.load(f)
    return paths, {"unit_rows": paths["vector_db_dir"], "scaled_rows": scaled}, golden


def test_oracle_matches_the_reference_script(env):
    paths, vec_dirs, golden = env
    assert sorted(golden) == ["scaled_rows", "unit_rows"]
    for label, vec_dir in vec_dirs.items():
        assert len(golden[label]) == 7  # broken / noschema have no report JSON in the clean directory: skipped (:277-283)
        for name, want in golden[label].items():
            rows = stubs.read_ixfi(Path(vec_dir) / f"{name}.faiss").rows
            got = oracle.cosine_stats_ref(oracle.cosine_matrix_ref(rows))
            assert want["chunks_count"] == rows.shape[0]
            for key in STAT_KEYS:
                assert abs(got[key] - want[key]) <= 1e-6, (label, name, key, got[key], want[key])
        # the rescaled copy holds an exactly parallel pair: the off-diagonal maximum clips at 1 (:98)
    assert all(v["max_similarity"] == 1.0 for v in golden["scaled_rows"].values())


@pytest.mark.gpu
def test_generator_mirror_matches_the_reference_script(env, tmp_path):
    from rag_challenge_2_b200 import cosine

    paths, vec_dirs, golden = env
    for label, vec_dir in vec_dirs.items():
        gen = cosine.SimilarityMatrixGenerator(vec_dir, paths["bm25_documents_dir"], tmp_path / label, pixels=16)
        stats = gen.process_all_documents()
        assert sorted(stats) == sorted(golden[label])
        with open(tmp_path / label / "statistics.json", encoding="utf-8") as f:
            assert json.load(f) == stats
        for name, want in golden[label].items():
            got = stats[name]
            for key in ("chunks_count", "company", "pages", "text_blocks", "tables"):
                assert got[key] == want[key]
            assert list(got) == ["chunks_count", *STAT_KEYS, "company", "pages", "text_blocks", "tables"]
            for key in ("avg_similarity", "std_similarity", "max_similarity", "min_similarity"):
                assert abs(got[key] - want[key]) <= 1e-5 * abs(want[key]) + 2e-6, (label, name, key, got[key], want[key])
            # report-sized inputs take the median from the values themselves (cosine.exact_median_small)
            assert abs(got["median_similarity"] - want["median_similarity"]) <= 2e-6
            assert (tmp_path / label / f"{name}_similarity_tiles.npy").exists()
