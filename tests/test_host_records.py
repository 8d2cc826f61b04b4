"""The CPython record helper (csrc/host_records.c) against the Python statement of the same loops
(`_RawHits._records_py`, reference src/retrieval.py:1142-1157 / :1552-1599 / :1639-1716) and against the builtin
`round` — host logic only, no GPU."""
import math
import struct

import numpy as np
import pytest

import rag_challenge_2_b200  # noqa: F401
from rag_challenge_2_b200 import retrieval, synth

helper = retrieval._b2r_records
pytestmark = pytest.mark.skipif(helper is None, reason="_b2r_records.so not built (run ./build.sh)")


def _same_float(a, b):
    return struct.pack("<d", a) == struct.pack("<d", b)


def test_round4_is_the_builtin_round():
    rng = np.random.default_rng(3)
    xs = [0.0, -0.0, 1.0, -1.0, 0.00005, -0.00005, 0.12345, 2.675, 1e-9, 123456.78905, float("inf"), float("-inf"),
          1e9, -1e9, 1e15, 1.7976931348623157e308, 5e-324, 0.30000000000000004, 0.99995, 0.999949999999]
    xs += rng.standard_normal(200_000).astype(np.float32).astype(np.float64).tolist()      # what a search returns
    xs += (rng.random(100_000) * 2 - 1).tolist()
    xs += (rng.standard_normal(20_000) * 1e3).tolist()
    # around every kind of tie: (k + 0.5) * 1e-4 and its floating-point neighbours
    for k in list(range(-50, 50)) + rng.integers(-10**7, 10**7, 3000).tolist():
        x = (k + 0.5) / 1e4
        for _ in range(3):
            xs.append(x)
            xs.append(math.nextafter(x, math.inf))
            xs.append(math.nextafter(x, -math.inf))
            x = math.nextafter(x, math.inf)
    for x in xs:
        got, want = helper.round4(x), round(x, 4)
        assert _same_float(got, want), (x.hex(), got, want)
    assert math.isnan(helper.round4(float("nan")))


def _rows(n_chunks, nq, k, seed):
    rng = np.random.default_rng(seed)
    I = np.stack([rng.choice(n_chunks, k, replace=False) for _ in range(nq)]).astype(np.int64)
    D = -np.sort(-rng.standard_normal((nq, k)).astype(np.float32), axis=1)
    return D, I


def _both(raw, query="问题 q"):
    logs_c, logs_py = [], []
    return raw.records(query, logs_c.append, helper=True), raw.records(query, logs_py.append, helper=False), logs_c, logs_py


@pytest.mark.parametrize("parent_pages", [False, True])
@pytest.mark.parametrize("lazy", [False, True])
def test_records_equal_the_python_loops(parent_pages, lazy):
    n = 5000
    rep = synth.synthetic_year_report("J2024", "c", n, None, chunks_per_page=4)
    chunks, pages = rep["document"]["content"]["chunks"], rep["_page_lookup"][1]
    if not lazy:
        chunks = [chunks[i] for i in range(n)]
        pages = retrieval._page_text_lookup([rep["document"]["content"]["pages"][i] for i in range((n + 3) // 4)])
    D, I = _rows(n, 9, 30, 1)
    I[2, :6] = [40, 41, 42, 43, 41 + 4, 40]   # page mode: several chunks of one page, in and out of order
    for row in range(9):
        raw = retrieval._RawHits("sha-1", chunks, pages, D, I, row, parent_pages, "J2024")
        got, want, logs_c, logs_py = _both(raw)
        assert got == want and not logs_c and not logs_py
        assert [list(r) for r in got] == [list(r) for r in want]          # key order (json.dumps of a record)
        for a, b in zip(got, want):
            assert _same_float(a["vector_similarity"], b["vector_similarity"])
            assert type(a["page"]) is type(b["page"]) and type(a["hit_count"]) is int
        if len(got) > 1:   # no list is shared between records
            assert got[0]["retrieval_sources"] is not got[1]["retrieval_sources"]
            assert got[0]["query_sources"] is not got[1]["query_sources"]
            got[0]["query_sources"].append("x")
            assert got[1]["query_sources"] == ["问题 q"]


@pytest.mark.parametrize("parent_pages", [False, True])
def test_a_failing_hit_ends_the_row_and_keeps_the_records_before_it(parent_pages):
    n = 200
    chunks = [{"page": i // 4 + 1, "text": f"chunk {i}"} for i in range(n)]
    pages = {p: {"page": p, "text": f"page {p}"} for p in range(1, n // 4 + 1)}
    D, I = _rows(n, 4, 10, 2)
    del pages[chunks[int(I[1, 3])]["page"]]                # a parent page is missing (reference :1145 fails the task)
    del chunks[int(I[2, 5])]["text"]                       # a damaged chunk record
    I[3, 7] = n + 5                                        # an id past the chunk list
    for row in range(4):
        raw = retrieval._RawHits("s", chunks, pages, D, I, row, parent_pages, "rep")
        got, want, logs_c, logs_py = _both(raw)
        assert got == want
        assert logs_c == logs_py, (logs_c, logs_py)
    raw = retrieval._RawHits("s", chunks, pages, D, I, 1, False, "rep")
    got, _, logs, _ = _both(raw)
    assert len(got) == 3 and len(logs) == 1 and "Vector search failed" in logs[0]


def test_rejects_rows_of_the_wrong_type():
    chunks = [{"page": 1, "text": "t"}]
    D, I = np.zeros((1, 1), np.float32), np.zeros((1, 1), np.int64)
    with pytest.raises(TypeError):
        helper.append_chunk_records([], D.astype(np.float64), I, 0, chunks, {1: {}}, "s", "q")
    with pytest.raises(TypeError):
        helper.append_chunk_records([], D, I.astype(np.int32), 0, chunks, {1: {}}, "s", "q")
    with pytest.raises(TypeError):
        helper.append_chunk_records([], D, I, 1, chunks, {1: {}}, "s", "q")


def test_direct_ok_uses_the_precomputed_flag_or_the_ids():
    D, I = np.zeros((3, 3), np.float32), np.array([[0, 1, 2], [1, 1, 2], [-1, 0, 1]], np.int64)
    assert retrieval._RawHits("s", [], {}, D, I, 0, False, "r").direct_ok()
    assert not retrieval._RawHits("s", [], {}, D, I, 1, False, "r").direct_ok()
    assert not retrieval._RawHits("s", [], {}, D, I, 2, False, "r").direct_ok()
    assert retrieval._RawHits("s", [], {}, D, I, 1, True, "r").direct_ok()
    assert not retrieval._RawHits("s", [], {}, D, I, 0, False, "r", ids_ok=False).direct_ok()


def test_other_result_dtypes_keep_the_python_loops():
    """An index object that returns float64 scores (not faiss's float32) must keep its own precision: round() of the
    float64 value, not of its float32 rounding."""
    chunks = [{"page": 1, "text": "t"}]
    pages = {1: {"page": 1, "text": "p"}}
    D = np.array([[0.123449999]], np.float64)            # float32(0.123449999) = 0.12345 -> would round to 0.1235
    I = np.array([[0]], np.int64)
    rec = retrieval._RawHits("s", chunks, pages, D, I, 0, False, "r").records("q", print)
    assert rec[0]["vector_similarity"] == round(0.123449999, 4) == 0.1234
    rec = retrieval._RawHits("s", chunks, pages, D.astype(np.float32), I.astype(np.int32), 0, False, "r").records("q", print)
    assert rec[0]["vector_similarity"] == round(float(np.float32(0.123449999)), 4)
