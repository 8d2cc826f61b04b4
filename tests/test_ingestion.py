"""Index builders (`BM25Ingestor`, `VectorDBIngestor`) against what the REAL src/ingestion.py produced over the same
chunked reports (tests/golden/make_ingestion_golden.py): embedding request sizes and error type on the CPU; on the
GPU the written `.faiss` rows byte for byte, the pickled BM25 state, and that the retrievers load the files."""
import json
import pickle

import pytest

from tests.golden import make_ingestion_golden as mk
from tests.golden import stubs


def golden():
    with open(mk.GOLDEN, encoding="utf-8") as f:
        return json.load(f)


def test_embedding_requests_and_errors_match_the_reference(tmp_path):
    from rag_challenge_2_b200.ingestion import VectorDBIngestor

    g = golden()
    assert sorted(g["vector"]) == sorted(g["bm25"]) and "BIG230" in g["vector"] and len(g["vector"]) == 9
    reports_dir = mk.build_reports_dir(tmp_path)
    qwen = mk.CountingQwen()
    ing = VectorDBIngestor(qwen=qwen, retry_wait=0.0)
    for path in sorted(reports_dir.glob("*.json")):
        with open(path, encoding="utf-8") as f:
            chunks = [c["text"] for c in json.load(f)["content"]["chunks"]]
        emb = ing._get_embeddings(chunks)
        assert len(emb) == len(chunks) and len(emb[0]) == stubs.DIM
    assert sorted(qwen.batches) == g["embedding_batches"]          # 230 chunks -> 100 + 100 + 30
    try:
        import tenacity  # noqa: F401
        expected = g["empty_text_error"]                           # RetryError in the authoring container
    except ImportError:
        expected = "ValueError"
    with pytest.raises(Exception) as info:
        ing._get_embeddings("   ")
    assert type(info.value).__name__ == expected


@pytest.mark.gpu
def test_ingestors_write_what_the_reference_writes(tmp_path):
    import rag_challenge_2_b200.retrieval as R
    from rag_challenge_2_b200.bm25 import load_bm25_pickle
    from rag_challenge_2_b200.ingestion import BM25Ingestor, VectorDBIngestor

    g = golden()
    reports_dir = mk.build_reports_dir(tmp_path / "db")
    vec_out, bm_out = tmp_path / "vec", tmp_path / "bm25"
    VectorDBIngestor(qwen=stubs.ScriptedQwen()).process_reports(reports_dir, vec_out)
    BM25Ingestor().process_reports(reports_dir, bm_out)
    assert sorted(p.stem for p in vec_out.glob("*.faiss")) == sorted(g["vector"])
    assert sorted(p.stem for p in bm_out.glob("*.pkl")) == sorted(g["bm25"])
    for name, want in g["vector"].items():
        assert mk.rows_digest(stubs.read_ixfi(vec_out / f"{name}.faiss").rows) == want, name
    for name, want in g["bm25"].items():
        obj = load_bm25_pickle(bm_out / f"{name}.pkl")
        got = mk.bm25_state(obj)
        for key in ("corpus_size", "doc_len", "k1", "b", "epsilon", "doc_freqs_sha256"):
            assert got[key] == want[key], (name, key)
        assert got["avgdl"] == want["avgdl"] and got["average_idf"] == pytest.approx(want["average_idf"], rel=1e-15)
        assert [t for t, _ in got["idf"]] == [t for t, _ in want["idf"]]
        assert [v for _, v in got["idf"]] == pytest.approx([v for _, v in want["idf"]], rel=1e-15)
        # raw pickle: the attribute set rank_bm25 pickles (plus our weight width)
        with open(bm_out / f"{name}.pkl", "rb") as f:
            raw = pickle.load(f)
        assert set(raw.__getstate__()) == {"corpus_size", "avgdl", "doc_freqs", "idf", "doc_len", "tokenizer", "k1", "b",
                                           "epsilon", "average_idf", "weight_bits"}
    # the written databases serve retrieval: same answers as over the fixture's own databases
    R.clear_resident_cache()
    a = R.VectorRetriever(vec_out, reports_dir, qwen=stubs.ScriptedQwen())
    res = a.retrieve_by_company_name(mk.corpus.COMPANY, mk.corpus.QUESTIONS[0], top_n=10, use_hyde=False,
                                     use_multi_query=False, selected_years=[2024])
    paths = mk.corpus.build(tmp_path / "fixture")
    b = R.VectorRetriever(paths["vector_db_dir"], paths["bm25_documents_dir"], qwen=stubs.ScriptedQwen())
    ref = b.retrieve_by_company_name(mk.corpus.COMPANY, mk.corpus.QUESTIONS[0], top_n=10, use_hyde=False,
                                     use_multi_query=False, selected_years=[2024])
    assert res["results"] == ref["results"]
    bm = R.BM25Retriever(bm_out, reports_dir)
    got = bm.retrieve_by_company_name(mk.corpus.COMPANY, "gross margin revenue 2024", top_n=5, selected_years=[2024])
    want = R.BM25Retriever(paths["bm25_db_dir"], paths["bm25_documents_dir"]).retrieve_by_company_name(
        mk.corpus.COMPANY, "gross margin revenue 2024", top_n=5, selected_years=[2024])
    assert got == want
    R.clear_resident_cache()


def test_ixfi_bytes_follow_the_field_order_of_faiss_index_write(tmp_path):
    """`.faiss` layout of an IndexFlatIP, derived field by field from faiss 1.9's `impl/index_write.cpp` (the package is not
    installable here, so this is the byte-level vector instead of a real file):
        write_index(): IndexFlat branch   -> WRITE1(fourcc("IxFI"))            4 bytes, the characters in order
        write_index_header(): WRITE1(d)            `int`                        4 bytes LE
                              WRITE1(ntotal)       `idx_t` = int64              8
                              WRITE1(dummy) x 2    `idx_t` = 1 << 20            8 + 8
                              WRITE1(is_trained)   `bool`                       1
                              WRITE1(metric_type)  `MetricType` enum (int), METRIC_INNER_PRODUCT = 0   4
                              (metric_arg only when metric_type > 1: not for IP)
        WRITEXBVECTOR(codes): size_t size = codes.size() / 4 (codes are the rows' bytes) -> number of float32   8
                              then the raw bytes                                  4 * ntotal * d
    i.e. a 45-byte prologue — the well-known size of flat faiss files, 45 + 4 * ntotal * d (a 1M x 128 IndexFlat is
    512,000,045 bytes).  The expected bytes below are spelled out by hand, not produced by the module's struct string."""
    import numpy as np

    from rag_challenge_2_b200 import flat_index

    class Rows:                                  # what write_index needs from an index: d, ntotal, reconstruct_n
        d, ntotal = 3, 2
        data = np.array([[1.0, -2.0, 0.5], [0.25, 3.0, -0.125]], dtype=np.float32)

        def reconstruct_n(self, lo, n):
            return self.data[lo: lo + n]

    path = tmp_path / "two_rows.faiss"
    flat_index.write_index(Rows(), path)
    blob = path.read_bytes()
    expected = (b"IxFI"
                + bytes([3, 0, 0, 0])                                 # d
                + bytes([2, 0, 0, 0, 0, 0, 0, 0])                     # ntotal
                + bytes([0, 0, 0x10, 0, 0, 0, 0, 0]) * 2              # the two dummies, 1 << 20
                + bytes([1])                                          # is_trained
                + bytes([0, 0, 0, 0])                                 # METRIC_INNER_PRODUCT
                + bytes([6, 0, 0, 0, 0, 0, 0, 0])                     # codes.size() / 4 = ntotal * d floats
                + bytes.fromhex("0000803f" "000000c0" "0000003f" "0000803e" "00004040" "000000be"))
    assert blob == expected and len(blob) == 45 + 4 * 2 * 3
    d, rows = flat_index.read_index_rows(path)
    assert d == 3 and np.array_equal(rows, Rows.data)
    assert np.array_equal(stubs.read_ixfi(path).rows, Rows.data)
    # an L2 flat index ("IxF2", metric 1) or a truncated file is refused, not misread
    with open(tmp_path / "l2.faiss", "wb") as f:
        f.write(b"IxF2" + expected[4:])
    with pytest.raises(RuntimeError):
        flat_index.read_index_rows(tmp_path / "l2.faiss")
    with open(tmp_path / "short.faiss", "wb") as f:
        f.write(expected[:-4])
    with pytest.raises(RuntimeError):
        flat_index.read_index_rows(tmp_path / "short.faiss")
