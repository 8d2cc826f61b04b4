"""GPU parity: K3 merge (float32 / float64), the dense+BM25 fusion kernel, and row-sharded search
(G shards on one GPU, gather emulated by stacking) against the oracle / the unsharded search."""
import numpy as np
import pytest

import oracle

pytestmark = pytest.mark.gpu


def _lists(rng, n_lists, nq, k_in, dtype, ties=False):
    D = rng.standard_normal((n_lists, nq, k_in)).astype(dtype)
    if ties:
        D = np.round(D, 1)
    D = -np.sort(-D, axis=2)
    I = rng.integers(0, 1 << 40, (n_lists, nq, k_in)).astype(np.int64)
    # ragged lists: some tails are empty (id -1)
    for l in range(n_lists):
        for q in range(nq):
            cut = int(rng.integers(0, k_in + 1)) if rng.random() < 0.3 else k_in
            I[l, q, cut:] = -1
    return D, I


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
@pytest.mark.parametrize("n_lists,nq,k_in,k_out", [(2, 3, 30, 30), (8, 5, 30, 30), (8, 2, 128, 100), (6, 4, 7, 20),
                                                    (1, 2, 5, 5), (8, 1, 1024, 1024), (40, 2, 300, 64)])
def test_merge_kernel(b2r, dtype, n_lists, nq, k_in, k_out):
    import torch

    from rag_challenge_2_b200.hybrid import merge_topk

    rng = np.random.default_rng(n_lists * 1000 + k_in)
    for ties in (False, True):
        D, I = _lists(rng, n_lists, nq, k_in, dtype, ties)
        pad = np.float32(-3.4028235e38) if dtype == np.float32 else -1.7976931348623157e308
        Dr, Ir = oracle.merge_lists_ref(D, I, k_out, pad)
        Dg, Ig = merge_topk(torch.from_numpy(D).cuda(), torch.from_numpy(I).cuda(), k_out)
        torch.cuda.synchronize()
        assert np.array_equal(Ig.cpu().numpy(), Ir)
        assert np.array_equal(Dg.cpu().numpy(), Dr)


def test_merge_host_entry(b2r):
    import ctypes as C

    from rag_challenge_2_b200 import _lib

    rng = np.random.default_rng(5)
    D, I = _lists(rng, 4, 3, 30, np.float32)
    Dr, Ir = oracle.merge_lists_ref(D, I, 30, np.float32(-3.4028235e38))
    Dg = np.empty((3, 30), np.float32)
    Ig = np.empty((3, 30), np.int64)
    _lib.check(_lib.load().b2r_topk_merge_host(_lib.ptr(D), _lib.ptr(I), 4, 3, 30, 30, _lib.ptr(Dg), _lib.ptr(Ig), 0))
    assert np.array_equal(Dg, Dr) and np.array_equal(Ig, Ir)


@pytest.mark.parametrize("k_dense,k_bm25,k_out", [(30, 30, 30), (30, 30, 10), (50, 10, 60), (1, 1, 1), (128, 128, 100)])
def test_fusion_kernel(b2r, k_dense, k_bm25, k_out):
    import torch

    from rag_challenge_2_b200.hybrid import hybrid_fuse

    rng = np.random.default_rng(k_dense + 7 * k_bm25)
    nq = 6
    D = -np.sort(-rng.random((nq, k_dense)).astype(np.float32), axis=1)
    S = -np.sort(-(rng.random((nq, k_bm25)) * 12), axis=1)
    I = np.stack([rng.choice(500, k_dense, replace=False) for _ in range(nq)]).astype(np.int64)
    J = np.stack([rng.choice(500, k_bm25, replace=False) for _ in range(nq)]).astype(np.int64)
    J[1, : min(k_bm25, k_dense) // 2] = I[1, : min(k_bm25, k_dense) // 2]   # heavy overlap
    S[2, :] = 3.25                                                           # constant BM25 list -> all 1.0
    S[3, :] = 0.0                                                            # all-zero BM25 list -> all 0.0
    J[4, k_bm25 // 2:] = -1                                                  # ragged
    I[5, k_dense // 3:] = -1
    Fr, Or, Mr = oracle.hybrid_fuse_ref(D, I, S, J, k_out)
    F, O, M = hybrid_fuse(torch.from_numpy(D).cuda(), torch.from_numpy(I).cuda(), torch.from_numpy(S).cuda(),
                          torch.from_numpy(J).cuda(), k_out, want_methods=True)
    torch.cuda.synchronize()
    assert np.array_equal(O.cpu().numpy(), Or)
    assert np.array_equal(F.cpu().numpy(), Fr)
    assert np.array_equal(M.cpu().numpy(), Mr)


def test_sharded_dense_and_hybrid_match_unsharded(b2r):
    """3 shards on one GPU, gather emulated by stacking: merged results == unsharded results, ids
    global, ties across shard boundaries resolved to the lower global id."""
    import torch

    from rag_challenge_2_b200 import hybrid, synth
    from rag_challenge_2_b200.bm25 import BM25Okapi

    n, d, V, k = 60_000, 1024, 5000, 30
    rows = synth.dense_rows_torch(0, n, d, "cuda")
    rows[40_000:40_050] = rows[100:150]          # duplicates across shards -> cross-shard ties
    q = synth.dense_queries_torch(5, d, "cuda", rows)
    q[0] = rows[120]                              # its best matches: row 120 and its copy 40020
    post = synth.token_postings_torch(0, n, V, "cuda", seed=77)
    idf, _ = synth.bm25_idf_from_df(post["df"].cpu().numpy(), n)
    idf_t = torch.from_numpy(idf).cuda()
    avgdl = post["total_len"] / n
    full_flat = b2r.FlatIPIndex.from_device_tensor(rows)
    full_bm = BM25Okapi.from_postings(post["indptr"], post["doc_ids"], post["tf"], post["doc_len"], idf_t, avgdl)
    whole = hybrid.HybridIndex(full_flat, full_bm, 0, world_size=1)
    queries = synth.token_queries_np(5, vocab=V, seed=78)
    offs = np.zeros(6, np.int32)
    np.cumsum([len(t) for t in queries], out=offs[1:])
    tids = torch.from_numpy(np.concatenate(queries).astype(np.int32)).cuda()
    toffs = torch.from_numpy(offs).cuda()
    D0, I0 = whole.search_dev(q, k)
    F0, O0 = whole.search_hybrid_dev(q, tids, toffs, k)
    assert I0[0, 0].item() == 120 and I0[0, 1].item() == 40020

    bounds = [0, 17_000, 41_000, n]
    dids, tf, dl = post["doc_ids"].long(), post["tf"], post["doc_len"]
    term_of = torch.repeat_interleave(torch.arange(V, device="cuda"), post["indptr"][1:] - post["indptr"][:-1])
    shards = []
    for lo, hi in zip(bounds[:-1], bounds[1:]):
        m = (dids >= lo) & (dids < hi)
        df = torch.bincount(term_of[m], minlength=V)
        indptr = torch.zeros(V + 1, dtype=torch.int64, device="cuda")
        indptr[1:] = torch.cumsum(df, 0)
        bm = BM25Okapi.from_postings(indptr, (dids[m] - lo).int(), tf[m], dl[lo:hi], idf_t, avgdl)
        fl = b2r.FlatIPIndex.from_device_tensor(rows[lo:hi])
        shards.append(hybrid.HybridIndex(fl, bm, lo, world_size=1))

    # emulate the SPMD all-gather: run every shard's local search, stack in rank order
    local = [(s.flat.search_dev(q, k, id_offset=s.offset), s.bm25.topk_dev(tids, toffs, k, id_offset=s.offset)) for s in shards]
    Dl = torch.stack([l[0][0] for l in local]); Il = torch.stack([l[0][1] for l in local])
    Sl = torch.stack([l[1][0] for l in local]); Jl = torch.stack([l[1][1] for l in local])
    D1, I1 = hybrid.merge_topk(Dl, Il, k)
    S1, J1 = hybrid.merge_topk(Sl, Jl, k)
    F1, O1 = hybrid.hybrid_fuse(D1, I1, S1, J1, k)
    torch.cuda.synchronize()
    assert torch.equal(I1, I0) and torch.equal(D1, D0)
    S0, J0 = whole.searcher.search_bm25(tids, toffs, k)
    assert torch.equal(J1, J0) and torch.equal(S1, S0)
    assert torch.equal(O1, O0) and torch.equal(F1, F0)
    # and the host-facing call agrees with the resident one
    Fh, Oh = whole.search_hybrid(q.cpu().numpy(), queries, k)
    assert np.array_equal(Oh, O0.cpu().numpy()) and np.array_equal(Fh, F0.cpu().numpy())
    # the one-call C step (b2r_hybrid_search_host) == the Python-driven host step, also one query at a time
    # (SIMT scan with BM25 launched beside it by programmatic dependent launch)
    Fp, Op = whole._search_hybrid_py(q.cpu().numpy(), queries, k)
    assert np.array_equal(Op, Oh) and np.array_equal(Fp, Fh)
    qh = q.cpu().numpy()
    for rep in range(3):
        for i in range(len(queries)):
            F1q, O1q = whole.search_hybrid(qh[i: i + 1], [queries[i]], k)
            assert np.array_equal(O1q[0], Oh[i])
            assert np.allclose(F1q[0], Fh[i], rtol=2e-6, atol=1e-7)   # K1 vs K2 summation order
    assert whole.last_h2d_bytes == 4 * 1024 + 8 + 4 * len(queries[-1]) and whole.last_d2h_bytes == 12 * k
    Fe, Oe = whole.search_hybrid(qh[:2], [[], queries[1]], k)          # a query without known tokens
    assert np.array_equal(Oe[1], Oh[1])


@pytest.mark.parametrize("world,k,step", [(1, 30, 1), (3, 30, 2), (8, 30, 5), (8, 100, 8), (4, 7, 3)])
def test_exchange_kernel_emulated(b2r, world, k, step):
    """The one-kernel exchange + merge + fusion, with the other ranks emulated on one GPU: their slots
    and flags are pre-filled in a plain buffer that stands for every rank's symmetric buffer (the
    kernel is launched once, as rank 0, so nothing waits on another launch)."""
    import torch

    from rag_challenge_2_b200 import _lib, hybrid

    nq, nq_max, k_max = 5, 8, 128
    rng = np.random.default_rng(world * 100 + k)
    D = -np.sort(-rng.random((world, nq, k)).astype(np.float32), axis=2)
    S = -np.sort(-(rng.random((world, nq, k)) * 9.0), axis=2)
    I = rng.integers(0, 5000, (world, nq, k)).astype(np.int64)
    J = rng.integers(0, 5000, (world, nq, k)).astype(np.int64)
    for r in range(world):                       # ids unique inside a rank's list; ragged tails
        for q in range(nq):
            I[r, q] = rng.choice(5000, k, replace=False) + 5000 * r
            J[r, q] = rng.choice(5000, k, replace=False) + 5000 * r
    I[0, 1, k // 2:] = -1
    J[world - 1, 2, 3:] = -1
    if world > 1:
        J[1, 0, :5] = I[0, 0, :5]                # overlap between dense and BM25 hits
        D[1, 3, 0] = D[0, 3, 0]                  # tie across ranks -> lower rank first
    nbytes = int(_lib.load().b2r_hybrid_exchange_bytes(nq_max, k_max, world))
    buf = np.zeros(nbytes, dtype=np.uint8)
    nk = nq_max * k_max
    slot = nk * 28
    parity = step & 1
    for r in range(1, world):
        base = (parity * world + r) * slot
        dD = buf[base: base + nk * 4].view(np.float32).reshape(nq_max, k_max)
        dI = buf[base + nk * 4: base + nk * 12].view(np.int64).reshape(nq_max, k_max)
        dS = buf[base + nk * 12: base + nk * 20].view(np.float64).reshape(nq_max, k_max)
        dJ = buf[base + nk * 20: base + nk * 28].view(np.int64).reshape(nq_max, k_max)
        dD[:nq, :k], dI[:nq, :k], dS[:nq, :k], dJ[:nq, :k] = D[r], I[r], S[r], J[r]
    flags = buf[2 * world * slot:].view(np.uint32).reshape(2, world, nq_max)
    flags[parity, 1:, :] = step
    dev = torch.from_numpy(buf).cuda()
    F, O = hybrid.exchange_fuse(torch.from_numpy(D[0]).cuda(), torch.from_numpy(I[0]).cuda(),
                                torch.from_numpy(S[0]).cuda(), torch.from_numpy(J[0]).cuda(),
                                [dev.data_ptr()] * world, 0, world, nq_max, k_max, step)
    torch.cuda.synchronize()
    Dm, Im = oracle.merge_lists_ref(D, I, k, np.float32(-3.4028235e38))
    Sm, Jm = oracle.merge_lists_ref(S, J, k, -1.7976931348623157e308)
    Fr, Or, _ = oracle.hybrid_fuse_ref(Dm, Im, Sm, Jm, k)
    assert np.array_equal(O.cpu().numpy(), Or)
    assert np.array_equal(F.cpu().numpy(), Fr)
    out_flags = dev.cpu().numpy()[2 * world * slot:].view(np.uint32).reshape(2, world, nq_max)
    assert np.all(out_flags[parity, 0, :nq] == step)       # rank 0 published its own flags


def test_chained_steps_equal_unchained(b2r):
    """A stream of hybrid steps chained by programmatic dependent launch (the scan of step s + 1 starts while the fusion
    kernel of step s finishes; searcher.pipeline) returns, step for step, what the same steps return one after the other —
    with outputs allocated per step from torch's caching allocator, i.e. with step s + 1's buffers aliasing step s's."""
    import torch

    from rag_challenge_2_b200 import hybrid, synth
    from rag_challenge_2_b200.bm25 import BM25Okapi

    n, d, V, k = 300_000, 1024, 5000, 30
    rows = synth.dense_rows_torch(0, n, d, "cuda")
    post = synth.token_postings_torch(0, n, V, "cuda", seed=91)
    idf, _ = synth.bm25_idf_from_df(post["df"].cpu().numpy(), n)
    bm = BM25Okapi.from_postings(post["indptr"], post["doc_ids"], post["tf"], post["doc_len"], torch.from_numpy(idf).cuda(),
                                 post["total_len"] / n)
    index = hybrid.HybridIndex(b2r.FlatIPIndex.from_device_tensor(rows), bm, 0, world_size=1)
    nsteps = 48
    q = synth.dense_queries_torch(nsteps, d, "cuda", rows, seed=92)
    toks = synth.token_queries_np(nsteps, vocab=V, seed=93)
    staged = [(q[i: i + 1].contiguous(), torch.tensor(toks[i], dtype=torch.int32, device="cuda"),
               torch.tensor([0, len(toks[i])], dtype=torch.int32, device="cuda")) for i in range(nsteps)]
    torch.cuda.synchronize()

    def run(pipeline):
        index.searcher.pipeline = pipeline
        out = []
        for qi, ti, oi in staged:
            F, O = index.search_hybrid_dev(qi, ti, oi, k)
            out.append((F.clone(), O.clone()))      # clones are enqueued behind the step, the originals are freed at once
        torch.cuda.synchronize()
        return out

    plain = run(False)
    for _ in range(3):
        chained = run(True)
        for (F0, O0), (F1, O1) in zip(plain, chained):
            assert torch.equal(O0, O1) and torch.equal(F0, F1)
    # two queries per step take the same kernels (K1 with NQ = 2)
    index.searcher.pipeline = True
    t2 = torch.tensor(toks[0] + toks[1], dtype=torch.int32, device="cuda")
    o2 = torch.tensor([0, len(toks[0]), len(toks[0]) + len(toks[1])], dtype=torch.int32, device="cuda")
    for _ in range(4):
        F2, O2 = index.search_hybrid_dev(q[:2].contiguous(), t2, o2, k)
    torch.cuda.synchronize()
    assert torch.equal(O2[0], plain[0][1][0]) and torch.equal(O2[1], plain[1][1][0])
    index.close()


@pytest.mark.parametrize("world,k,step,nq", [(1, 30, 1, 1), (3, 30, 2, 2), (8, 30, 5, 1), (8, 100, 8, 2), (4, 7, 3, 2)])
def test_fused_tail_step_emulated_ranks(b2r, world, k, step, nq):
    """b2r_hybrid_step_dev: scan + BM25 whose last CTA per query fuses (world 1) or exchanges + merges + fuses in place
    (world > 1, the other ranks emulated as in test_exchange_kernel_emulated: their slots and flags pre-filled in one
    buffer that stands for every rank's symmetric buffer).  Checked against the oracle applied to this rank's real hits
    and the emulated ranks' lists; D / I / S / J must equal the separate searches."""
    import ctypes as C

    import torch

    from rag_challenge_2_b200 import _lib, synth
    from rag_challenge_2_b200.bm25 import BM25Okapi

    n, d, V = 40_000, 1024, 3000
    nq_max, k_max = 8, 128
    rows = synth.dense_rows_torch(0, n, d, "cuda", seed=300 + world)
    flat = b2r.FlatIPIndex.from_device_tensor(rows)
    post = synth.token_postings_torch(0, n, V, "cuda", seed=301)
    idf, _ = synth.bm25_idf_from_df(post["df"].cpu().numpy(), n)
    bm = BM25Okapi.from_postings(post["indptr"], post["doc_ids"], post["tf"], post["doc_len"], torch.from_numpy(idf).cuda(),
                                 post["total_len"] / n)
    q = synth.dense_queries_torch(nq, d, "cuda", rows, seed=302)
    toks = synth.token_queries_np(nq, vocab=V, seed=303)
    offs = np.zeros(nq + 1, np.int32)
    np.cumsum([len(t) for t in toks], out=offs[1:])
    tids = torch.from_numpy(np.concatenate(toks).astype(np.int32)).cuda()
    toffs = torch.from_numpy(offs).cuda()
    D0, I0 = flat.search_dev(q, k)
    S0, J0 = bm.topk_dev(tids, toffs, k)
    torch.cuda.synchronize()
    rng = np.random.default_rng(world * 100 + k)
    D = np.zeros((world, nq, k), np.float32)
    S = np.zeros((world, nq, k))
    I = np.zeros((world, nq, k), np.int64)
    J = np.zeros((world, nq, k), np.int64)
    D[0], I[0], S[0], J[0] = D0.cpu().numpy(), I0.cpu().numpy(), S0.cpu().numpy(), J0.cpu().numpy()
    for r in range(1, world):
        D[r] = -np.sort(-(rng.random((nq, k)).astype(np.float32) * float(D[0].max()) * 1.2), axis=1)
        S[r] = -np.sort(-(rng.random((nq, k)) * float(S[0].max()) * 1.2), axis=1)
        for qi in range(nq):
            I[r, qi] = rng.choice(n, k, replace=False) + n * r
            J[r, qi] = rng.choice(n, k, replace=False) + n * r
    if world > 1:
        J[1, 0, :3] = I[0, 0, :3]                # found by both retrievers, on different ranks
        I[world - 1, 0, k // 2:] = -1            # a ragged list
    nbytes = int(_lib.load().b2r_hybrid_exchange_bytes(nq_max, k_max, world))
    buf = np.zeros(nbytes, dtype=np.uint8)
    nk = nq_max * k_max
    slot = nk * 28
    parity = step & 1
    for r in range(1, world):
        base = (parity * world + r) * slot
        buf[base: base + nk * 4].view(np.float32).reshape(nq_max, k_max)[:nq, :k] = D[r]
        buf[base + nk * 4: base + nk * 12].view(np.int64).reshape(nq_max, k_max)[:nq, :k] = I[r]
        buf[base + nk * 12: base + nk * 20].view(np.float64).reshape(nq_max, k_max)[:nq, :k] = S[r]
        buf[base + nk * 20: base + nk * 28].view(np.int64).reshape(nq_max, k_max)[:nq, :k] = J[r]
    buf[2 * world * slot:].view(np.uint32).reshape(2, world, nq_max)[parity, 1:, :] = step
    dev = torch.from_numpy(buf).cuda()
    status = torch.zeros(1, dtype=torch.int32, device="cuda")
    out = {name: torch.empty((nq, k), dtype=dt, device="cuda") for name, dt in
           (("D", torch.float32), ("I", torch.int64), ("S", torch.float64), ("J", torch.int64), ("F", torch.float32), ("O", torch.int64))}
    arr = (C.c_void_p * world)(*[C.c_void_p(dev.data_ptr())] * world)
    n_launch = C.c_int(0)
    for flags in (0, 1):                          # plain, and launched as a chained step
        for t in out.values():
            t.zero_()
        if world > 1:                             # rank 0's own flags of this step are cleared between the two runs
            dev[2 * world * slot:].view(torch.int32).view(2, world, nq_max)[parity, 0, :] = 0
        torch.cuda.synchronize()
        _lib.check(_lib.load().b2r_hybrid_step_dev(
            flat._h, bm._h, C.c_void_p(q.data_ptr()), nq, C.c_void_p(tids.data_ptr()), C.c_void_p(toffs.data_ptr()), k, 0,
            arr if world > 1 else None, 0, world, nq_max, k_max, step, C.c_void_p(status.data_ptr()),
            *[C.c_void_p(out[name].data_ptr()) for name in ("D", "I", "S", "J", "F", "O")],
            C.c_void_p(torch.cuda.current_stream().cuda_stream), flags, C.byref(n_launch)))
        torch.cuda.synchronize()
        assert n_launch.value == 2 and int(status.item()) == 0
        assert torch.equal(out["D"], D0) and torch.equal(out["I"], I0) and torch.equal(out["S"], S0) and torch.equal(out["J"], J0)
        Dm, Im = oracle.merge_lists_ref(D, I, k, np.float32(-3.4028235e38))
        Sm, Jm = oracle.merge_lists_ref(S, J, k, -1.7976931348623157e308)
        Fr, Or, _ = oracle.hybrid_fuse_ref(Dm, Im, Sm, Jm, k)
        assert np.array_equal(out["O"].cpu().numpy(), Or)
        assert np.array_equal(out["F"].cpu().numpy(), Fr)
    # the three-kernel path gives the same answer
    _lib.tunable_set("hybrid.no_fused_tail", 1)
    try:
        if world > 1:
            dev[2 * world * slot:].view(torch.int32).view(2, world, nq_max)[parity, 0, :] = 0
        F3, O3 = torch.empty_like(out["F"]), torch.empty_like(out["O"])
        _lib.check(_lib.load().b2r_hybrid_step_dev(
            flat._h, bm._h, C.c_void_p(q.data_ptr()), nq, C.c_void_p(tids.data_ptr()), C.c_void_p(toffs.data_ptr()), k, 0,
            arr if world > 1 else None, 0, world, nq_max, k_max, step, C.c_void_p(status.data_ptr()),
            *[C.c_void_p(out[name].data_ptr()) for name in ("D", "I", "S", "J")], C.c_void_p(F3.data_ptr()),
            C.c_void_p(O3.data_ptr()), C.c_void_p(torch.cuda.current_stream().cuda_stream), 0, C.byref(n_launch)))
        torch.cuda.synchronize()
    finally:
        _lib.tunable_set("hybrid.no_fused_tail", 0)
    assert n_launch.value >= 3 and torch.equal(F3, out["F"]) and torch.equal(O3, out["O"])


def test_submit_collect_two_in_flight(b2r):
    """search_hybrid_submit / _collect: the next question is submitted while the previous one is on the GPU; answers equal
    the blocking call's, a third outstanding submission is refused, tickets are collected in order."""
    import torch

    from rag_challenge_2_b200 import _lib, hybrid, synth
    from rag_challenge_2_b200.bm25 import BM25Okapi

    n, d, V, k = 120_000, 1024, 4000, 30
    rows = synth.dense_rows_torch(0, n, d, "cuda", seed=401)
    post = synth.token_postings_torch(0, n, V, "cuda", seed=402)
    idf, _ = synth.bm25_idf_from_df(post["df"].cpu().numpy(), n)
    bm = BM25Okapi.from_postings(post["indptr"], post["doc_ids"], post["tf"], post["doc_len"], torch.from_numpy(idf).cuda(),
                                 post["total_len"] / n)
    index = hybrid.HybridIndex(b2r.FlatIPIndex.from_device_tensor(rows), bm, 0, world_size=1)
    q = synth.dense_queries_torch(12, d, "cuda", rows, seed=403).cpu().numpy()
    toks = synth.token_queries_np(12, vocab=V, seed=404)
    want = [index.search_hybrid(q[i: i + 1], [toks[i]], k) for i in range(12)]
    pending = index.search_hybrid_submit(q[0:1], [toks[0]], k)
    got = []
    for i in range(1, 12):
        nxt = index.search_hybrid_submit(q[i: i + 1], [toks[i]], k)
        got.append(index.search_hybrid_collect(pending))
        pending = nxt
    got.append(index.search_hybrid_collect(pending))
    for (F, O), (Fw, Ow) in zip(got, want):
        assert np.array_equal(O, Ow) and np.array_equal(F, Fw)
    t0 = index.search_hybrid_submit(q[0:1], [toks[0]], k)
    t1 = index.search_hybrid_submit(q[1:2], [toks[1]], k)
    with pytest.raises(_lib.B2RError):
        index.search_hybrid_submit(q[2:3], [toks[2]], k)          # both slots in flight
    F0, O0 = index.search_hybrid_collect(t0)
    F1, O1 = index.search_hybrid_collect(t1)
    assert np.array_equal(O0, want[0][1]) and np.array_equal(O1, want[1][1])
    with pytest.raises(_lib.B2RError):
        index.search_hybrid_collect(t0)                            # nothing outstanding under that ticket
    index.close()
