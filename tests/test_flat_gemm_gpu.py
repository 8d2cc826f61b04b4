"""GPU parity: K2, the tcgen05 (3xTF32) batched scan, through the C ABI, vs the float64 oracle.
Same bar as the SIMT path: ids equal up to swaps inside a 1e-5-relative band, scores 1e-5 relative."""
import numpy as np
import pytest

import oracle

pytestmark = pytest.mark.gpu


def unit_rows(n, d, seed):
    rng = np.random.default_rng(seed)
    x = rng.standard_normal((n, d), dtype=np.float32)
    x /= np.linalg.norm(x, axis=1, keepdims=True)
    return x


def queries_for(xb, nq, seed):
    rng = np.random.default_rng(seed)
    d = xb.shape[1]
    q = rng.standard_normal((nq, d), dtype=np.float32)
    for i in range(0, nq, 2):
        j = int(rng.integers(0, xb.shape[0]))
        q[i] = xb[j] + 0.1 * rng.standard_normal(d, dtype=np.float32) / np.sqrt(d)
    q /= np.linalg.norm(q, axis=1, keepdims=True)
    return q


def check(index, xb, xq, k):
    D, I = index.search(xq, k)
    S = oracle.flat_ip_scores_ref(xb, xq, np.float64)
    oracle.assert_topk_equivalent(D, I, S, k, rtol=1e-5, what=f"gemm n={xb.shape[0]} d={xb.shape[1]} nq={xq.shape[0]} k={k}")
    return D, I


@pytest.mark.parametrize("nq", [5, 16, 17, 32, 33, 64, 65, 130])
def test_batches_d1024(b2r, nq):
    xb = unit_rows(30011, 1024, 41)
    index = b2r.FlatIPIndex(1024)
    index.add(xb)
    check(index, xb, queries_for(xb, nq, 42 + nq), 30)
    assert index.last_launch_count() >= 3      # split + tcgen05 scan + select (per <= 64 queries)


@pytest.mark.parametrize("n", [128, 129, 255, 1000, 18945])
@pytest.mark.parametrize("k", [1, 10, 64, 128])
def test_tile_edges_and_k(b2r, n, k):
    xb = unit_rows(n, 1024, 43)
    index = b2r.FlatIPIndex(1024)
    index.add(xb)
    check(index, xb, queries_for(xb, 12, 44), k)


@pytest.mark.parametrize("d", [32, 96, 256, 768, 1536, 2048])
def test_other_widths(b2r, d):
    xb = unit_rows(9001, d, 45)
    index = b2r.FlatIPIndex(d)
    index.add(xb)
    check(index, xb, queries_for(xb, 20, 46), 30)


def test_unnormalised_large_magnitudes(b2r):
    """3xTF32 error is relative to |a||b|: check rows / queries far from unit norm."""
    rng = np.random.default_rng(47)
    xb = (rng.standard_normal((20000, 1024)) * rng.uniform(0.01, 50.0, (20000, 1))).astype(np.float32)
    xq = (rng.standard_normal((24, 1024)) * 7.0).astype(np.float32)
    index = b2r.FlatIPIndex(1024)
    index.add(xb)
    D, I = index.search(xq, 30)
    S = oracle.flat_ip_scores_ref(xb, xq, np.float64)
    # tolerance relative to the score scale of each query (|q| * max|x|), as the SIMT path
    scale = np.abs(S).max(axis=1, keepdims=True)
    got = np.take_along_axis(S, I, 1)
    assert np.all(np.abs(D - got) <= 1e-5 * scale)
    kth = np.sort(S, axis=1)[:, -30][:, None]
    assert np.all(got >= kth - 1e-5 * scale)


def test_ties_and_matches_simt_path(b2r):
    from rag_challenge_2_b200 import _lib

    base = unit_rows(64, 1024, 48)
    xb = np.concatenate([base] * 80, axis=0)          # heavy exact ties: row i == row i + 64 j
    index = b2r.FlatIPIndex(1024)
    index.add(xb)
    xq = base[:8].copy()
    D, I = index.search(xq, 30)
    for qi in range(8):
        top = I[qi][D[qi] == D[qi, 0]]
        assert np.array_equal(top, qi + 64 * np.arange(len(top)))   # equal scores: ascending id
    xb2 = unit_rows(50000, 1024, 49)
    index2 = b2r.FlatIPIndex(1024)
    index2.add(xb2)
    q2 = queries_for(xb2, 8, 50)
    D1, I1 = check(index2, xb2, q2, 30)
    _lib.tunable_set("flat.no_gemm", 1)
    try:
        D2, I2 = check(index2, xb2, q2, 30)
    finally:
        _lib.tunable_set("flat.no_gemm", 0)
    assert np.array_equal(I1, I2)
    assert np.allclose(D1, D2, rtol=1e-5, atol=1e-6)


def test_3xtf32_mode(b2r):
    """gemm.mode = 3: the 3xTF32 kernel alone (the redo path of the default mode, and k > 64)."""
    from rag_challenge_2_b200 import _lib

    xb = unit_rows(25000, 1024, 51)
    index = b2r.FlatIPIndex(1024)
    index.add(xb)
    _lib.tunable_set("gemm.mode", 3)
    try:
        for nq in (7, 40, 64):
            check(index, xb, queries_for(xb, nq, 52 + nq), 30)
    finally:
        _lib.tunable_set("gemm.mode", 0)


def test_unprovable_queries_are_redone(b2r):
    """Thousands of rows within 1e-4 of each other around the k-th score: the tf32 candidate set
    cannot be proven complete, so those queries must go through the 3xTF32 redo pass and still
    satisfy the parity bar; with verification disabled the approximate pass alone may miss rows."""
    rng = np.random.default_rng(53)
    base = unit_rows(8, 1024, 54)
    reps = 4000
    xb = np.repeat(base, reps, axis=0) + 1e-4 * rng.standard_normal((8 * reps, 1024)).astype(np.float32) / 32.0
    xb = np.concatenate([xb, unit_rows(8000, 1024, 55)]).astype(np.float32)
    xq = np.concatenate([base, unit_rows(4, 1024, 56)])          # 8 "crowded" queries + 4 ordinary ones
    index = b2r.FlatIPIndex(1024)
    index.add(xb)
    check(index, xb, xq, 30)


def test_exact_scores_equal_simt_scores(b2r):
    """Rescoring recomputes the scores in fp32 on the SIMT pipes: for well-separated data the
    tensor-core path returns the same ids as the nq <= 4 path and scores equal to the last ulp."""
    from rag_challenge_2_b200 import _lib

    xb = unit_rows(40000, 1024, 57)
    xq = queries_for(xb, 16, 58)
    index = b2r.FlatIPIndex(1024)
    index.add(xb)
    D1, I1 = index.search(xq, 30)
    _lib.tunable_set("flat.no_gemm", 1)
    try:
        D2, I2 = index.search(xq, 30)
    finally:
        _lib.tunable_set("flat.no_gemm", 0)
    assert np.array_equal(I1, I2)
    assert np.allclose(D1, D2, rtol=1e-6, atol=1e-7)   # both fp32 SIMT sums; only the reduction order differs


@pytest.mark.parametrize("nq", [100, 128, 129, 300])
def test_wide_threshold_pass(b2r, nq):
    """n >= 16384 rows: the candidate pass is the threshold kernel (flat_gemm_wide.cu), up to 128 queries
    per pass; same parity bar."""
    xb = unit_rows(70001, 1024, 61)
    index = b2r.FlatIPIndex(1024)
    index.add(xb)
    check(index, xb, queries_for(xb, nq, 62 + nq), 30)


def test_wide_equals_list_scan(b2r):
    """Threshold candidate pass vs the list-keeping candidate pass (gemm.no_wide): identical answers."""
    from rag_challenge_2_b200 import _lib

    xb = unit_rows(60000, 1024, 63)
    xq = queries_for(xb, 64, 64)
    index = b2r.FlatIPIndex(1024)
    index.add(xb)
    D1, I1 = check(index, xb, xq, 30)
    _lib.tunable_set("gemm.no_wide", 1)
    try:
        D2, I2 = check(index, xb, xq, 30)
    finally:
        _lib.tunable_set("gemm.no_wide", 0)
    assert np.array_equal(I1, I2)
    assert np.array_equal(D1, D2)   # both rescore the same candidates with the same fp32 arithmetic


def test_wide_adversarial_order_overflows_to_redo(b2r):
    """Rows sorted by ascending score for the queries: the sampled thresholds (taken from the first tiles of
    every CTA) are far too low, the candidate lists overflow, and the flagged queries must be answered by
    the robust 3xTF32 pass."""
    xb = unit_rows(40000, 1024, 65)
    q = unit_rows(6, 1024, 66)
    order = np.argsort(xb @ q[0])
    xb = np.ascontiguousarray(xb[order])
    xq = np.concatenate([q, queries_for(xb, 10, 67)])
    index = b2r.FlatIPIndex(1024)
    index.add(xb)
    check(index, xb, xq, 30)


@pytest.mark.parametrize("d", [64, 512, 4096])
def test_wide_other_widths(b2r, d):
    xb = unit_rows(20000, d, 68)
    index = b2r.FlatIPIndex(d)
    index.add(xb)
    check(index, xb, queries_for(xb, 70, 69), 10)


@pytest.mark.parametrize("nq,sub", [(129, 0), (256, 0), (257, 0), (600, 0), (200, 2)])
def test_cta_pair_pass(b2r, nq, sub):
    """More than 128 queries: the CTA-pair kernel (tcgen05.mma.cta_group::2, 256 queries per pass), in both of its
    TMEM layouts (one sub-tile double-buffered / two sub-tiles single-buffered); same parity bar, and the same answers as
    with the pair kernel disabled."""
    from rag_challenge_2_b200 import _lib

    xb = unit_rows(90001, 1024, 81)
    index = b2r.FlatIPIndex(1024)
    index.add(xb)
    xq = queries_for(xb, nq, 82 + nq)
    _lib.tunable_set("gemm.wide_sub", sub)
    try:
        D1, I1 = check(index, xb, xq, 30)
    finally:
        _lib.tunable_set("gemm.wide_sub", 0)
    _lib.tunable_set("gemm.no_pair", 1)
    try:
        D2, I2 = index.search(xq, 30)
    finally:
        _lib.tunable_set("gemm.no_pair", 0)
    assert np.array_equal(I1, I2) and np.array_equal(D1, D2)


def test_cta_pair_other_widths_and_overflow(b2r):
    for d in (64, 2048):
        xb = unit_rows(30000, d, 83)
        index = b2r.FlatIPIndex(d)
        index.add(xb)
        check(index, xb, queries_for(xb, 150, 84), 10)
    # adversarial row order: thresholds sampled from the first tiles are too low, lists overflow, redo pass answers
    xb = unit_rows(40000, 1024, 85)
    q = unit_rows(3, 1024, 86)
    xb = np.ascontiguousarray(xb[np.argsort(xb @ q[0])])
    index = b2r.FlatIPIndex(1024)
    index.add(xb)
    check(index, xb, np.concatenate([q, queries_for(xb, 140, 87)]), 30)


@pytest.mark.parametrize("nq,d", [(256, 1024), (300, 1024), (200, 192)])
def test_bf16_candidate_pass(b2r, nq, d):
    """Opt-in bf16 copy of the rows for the candidate pass of large batches: the answers stay exact (fp32 rescoring,
    completeness proof with the bf16 bound, 3xTF32 redo) and equal the tf32 path's."""
    from rag_challenge_2_b200 import _lib

    xb = unit_rows(60001, d, 91)
    index = b2r.FlatIPIndex(d)
    index.add(xb)
    xq = queries_for(xb, nq, 92 + nq)
    D0, I0 = check(index, xb, xq, 30)
    index.enable_bf16_candidates()
    D1, I1 = check(index, xb, xq, 30)
    assert np.array_equal(I0, I1) and np.array_equal(D0, D1)
    index.add(unit_rows(5000, d, 93))                      # growing the index rebuilds the copy
    xb2 = np.concatenate([xb, unit_rows(5000, d, 93)])
    check(index, xb2, xq, 30)
    # crowded scores: the bf16 bound cannot prove completeness, the redo pass must answer
    base = unit_rows(4, d, 94)
    rng = np.random.default_rng(95)
    crowd = np.repeat(base, 3000, axis=0) + 1e-4 * rng.standard_normal((12000, d)).astype(np.float32) / np.sqrt(d)
    xb3 = np.concatenate([crowd, unit_rows(20000, d, 96)]).astype(np.float32)
    idx3 = b2r.FlatIPIndex(d)
    idx3.add(xb3)
    idx3.enable_bf16_candidates()
    check(idx3, xb3, np.concatenate([base, queries_for(xb3, 140, 97)]), 30)
    index.enable_bf16_candidates(False)
    D2, I2 = index.search(xq, 30)
    _lib.tunable_set("gemm.no_pair", 0)
    assert I2.shape == (nq, 30)
