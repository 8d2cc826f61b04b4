"""CPU, world_size 2, gloo: the SPMD plumbing of ShardedSearcher (shard bounds, global ids, the
packed all-gather, list-major merge order, fusion) with oracle stand-ins for the CUDA kernels."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import oracle


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _cpu_merge(D_lists, I_lists, k):
    pad = np.float32(-3.4028235e38) if D_lists.dtype == torch.float32 else -1.7976931348623157e308
    D, I = oracle.merge_lists_ref(D_lists.numpy(), I_lists.numpy(), k, pad)
    return torch.from_numpy(D), torch.from_numpy(I)


def _cpu_fuse(D, I, S, J, k):
    F, O, _ = oracle.hybrid_fuse_ref(D.numpy(), I.numpy(), S.numpy(), J.numpy(), k)
    return torch.from_numpy(F), torch.from_numpy(O)


def _worker(rank, world, port, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from rag_challenge_2_b200.hybrid import ShardedSearcher, shard_bounds

    n, d, k, V = 5000, 64, 30, 300
    rng = np.random.default_rng(0)
    xb = rng.standard_normal((n, d)).astype(np.float32)
    xb[3000:3010] = xb[10:20]      # ties across the shard boundary
    xq = np.concatenate([xb[[12, 4000]], rng.standard_normal((3, d)).astype(np.float32)])
    docs = [[f"t{int(t)}" for t in rng.integers(0, V, int(rng.integers(5, 30)))] for _ in range(n)]
    bm_full = oracle.BM25OkapiRef(docs)
    queries = [[f"t{int(t)}" for t in rng.integers(0, V, 4)] for _ in range(5)]
    bounds = shard_bounds(n, world, align=128)
    lo, hi = bounds[rank], bounds[rank + 1]

    def local_dense(q, kk):
        D, I = oracle.flat_ip_search_ref(xb[lo:hi], q.numpy(), kk)
        return torch.from_numpy(D), torch.from_numpy(np.where(I >= 0, I + lo, -1))

    def local_bm25(tids, toffs, kk):
        S = np.full((len(queries), kk), -1.7976931348623157e308)
        J = np.full((len(queries), kk), -1, dtype=np.int64)
        for qi, q in enumerate(queries):
            sc = bm_full.get_scores(q)[lo:hi]          # global idf / avgdl, local documents
            top = oracle.bm25_topn_ref(sc, kk)
            S[qi, :len(top)] = sc[top]
            J[qi, :len(top)] = np.array(top) + lo
        return torch.from_numpy(S), torch.from_numpy(J)

    s = ShardedSearcher(local_dense, local_bm25, merge=_cpu_merge, fuse=_cpu_fuse)
    assert s.world_size == world
    D, I = s.search_dense(torch.from_numpy(xq), k)
    assert s.collectives == 1
    S, J = s.search_bm25(None, None, k)
    F, O = s.search_hybrid(torch.from_numpy(xq), None, None, k)
    assert s.collectives == 1      # dense and BM25 lists travel in one all-gather
    np.savez(os.path.join(out_dir, f"r{rank}.npz"), D=D.numpy(), I=I.numpy(), S=S.numpy(), J=J.numpy(), F=F.numpy(),
             O=O.numpy(), bounds=np.array(bounds))
    dist.barrier()
    dist.destroy_process_group()


def test_sharded_search_world2_gloo(tmp_path):
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    r0, r1 = np.load(tmp_path / "r0.npz"), np.load(tmp_path / "r1.npz")
    for key in ("D", "I", "S", "J", "F", "O"):
        assert np.array_equal(r0[key], r1[key]), f"ranks disagree on {key}"
    assert r0["bounds"].tolist() == [0, 2560, 5000]
    # against the unsharded oracle
    n, d, k, V = 5000, 64, 30, 300
    rng = np.random.default_rng(0)
    xb = rng.standard_normal((n, d)).astype(np.float32)
    xb[3000:3010] = xb[10:20]
    xq = np.concatenate([xb[[12, 4000]], rng.standard_normal((3, d)).astype(np.float32)])
    docs = [[f"t{int(t)}" for t in rng.integers(0, V, int(rng.integers(5, 30)))] for _ in range(n)]
    bm_full = oracle.BM25OkapiRef(docs)
    queries = [[f"t{int(t)}" for t in rng.integers(0, V, 4)] for _ in range(5)]
    Dr, Ir = oracle.flat_ip_search_ref(xb, xq, k)
    assert np.array_equal(r0["I"], Ir) and np.array_equal(r0["D"], Dr)
    assert r0["I"][0, 0] == 12 and r0["I"][0, 1] == 3002      # tie across shards: lower global id first
    Sr = np.stack([bm_full.get_scores(q) for q in queries])
    Jr = np.array([oracle.bm25_topn_ref(s, k) for s in Sr])
    assert np.array_equal(r0["J"], Jr)
    assert np.array_equal(r0["S"], np.take_along_axis(Sr, Jr, 1))
    Fr, Or, _ = oracle.hybrid_fuse_ref(Dr, Ir, np.take_along_axis(Sr, Jr, 1), Jr, k)
    assert np.array_equal(r0["O"], Or) and np.array_equal(r0["F"], Fr)


def test_shard_bounds():
    from rag_challenge_2_b200.hybrid import shard_bounds

    b = shard_bounds(10_000_000, 8, 128)                      # the bench's sharding: balanced to one 128-row block
    sizes = [hi - lo for lo, hi in zip(b[:-1], b[1:])]
    assert b[0] == 0 and b[-1] == 10_000_000 and all(x % 128 == 0 for x in b[:-1])
    assert max(sizes) - min(sizes) <= 128 and max(sizes) == 1_250_048 and sizes[-1] == min(sizes)
    assert shard_bounds(10_000_000, 8, 1 << 14)[:4] == [0, 1261568, 2523136, 3784704]   # 611 blocks: 3 ranks get 77, 5 get 76
    assert shard_bounds(10, 4) == [0, 3, 6, 8, 10]
    assert shard_bounds(5, 8) == [0, 1, 2, 3, 4, 5, 5, 5, 5]
    assert shard_bounds(1000, 1, 128) == [0, 1000]
    for n, w, a in ((1000, 3, 128), (129, 2, 128), (7, 3, 4)):
        bb = shard_bounds(n, w, a)
        assert bb[0] == 0 and bb[-1] == n and all(x <= y for x, y in zip(bb[:-1], bb[1:]))


def _token_corpus(n=1200, V=300, seed=5):
    rng = np.random.default_rng(seed)
    docs = [[int(t) for t in rng.integers(0, V, int(rng.integers(3, 25)))] for _ in range(n)]
    docs[7] = [V - 1] * 4                     # a term that first appears early and only on rank 0
    offs = np.zeros(n + 1, np.int64)
    np.cumsum([len(d) for d in docs], out=offs[1:])
    return docs, np.concatenate([np.asarray(d, np.int32) for d in docs]), offs


def _stats_worker(rank, world, port, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from rag_challenge_2_b200.bm25 import token_stream_statistics

    docs, tok, offs = _token_corpus()
    cut = [0, 700, len(docs)]
    lo, hi = cut[rank], cut[rank + 1]
    st = token_stream_statistics(tok[offs[lo]: offs[hi]], offs[lo: hi + 1] - offs[lo], 300, device=torch.device("cpu"),
                                 group=dist.group.WORLD)
    np.savez(os.path.join(out_dir, f"s{rank}.npz"), idf=st["idf"], avgdl=st["avgdl"], average_idf=st["average_idf"],
             corpus_size=st["corpus_size"], indptr=st["indptr"].numpy(), doc_ids=st["doc_ids"].numpy(), tf=st["tf"].numpy())
    dist.barrier()
    dist.destroy_process_group()


def test_token_stream_statistics_sharded_world2_gloo(tmp_path):
    """BM25 index build from a token stream: a 2-way document shard (df, token / document counts and the terms'
    first-seen order all-reduced) must produce the corpus-global idf / avgdl of the unsharded build, which in turn
    equals rank_bm25's dictionary arithmetic (corpus_statistics) bit for bit."""
    from rag_challenge_2_b200.bm25 import corpus_statistics, token_stream_statistics

    mp.spawn(_stats_worker, args=(2, _free_port(), str(tmp_path)), nprocs=2, join=True)
    docs, tok, offs = _token_corpus()
    whole = token_stream_statistics(tok, offs, 300, device=torch.device("cpu"))
    nd = {}
    for d in docs:
        for w in dict.fromkeys(d):
            nd[w] = nd.get(w, 0) + 1
    idf, avg = corpus_statistics(nd, len(docs))
    assert whole["average_idf"] == avg and whole["avgdl"] == sum(map(len, docs)) / len(docs)
    assert all(whole["idf"][w] == v for w, v in idf.items())
    s0, s1 = np.load(tmp_path / "s0.npz"), np.load(tmp_path / "s1.npz")
    for s in (s0, s1):
        assert np.array_equal(s["idf"], whole["idf"]) and float(s["avgdl"]) == whole["avgdl"]
        assert float(s["average_idf"]) == whole["average_idf"] and int(s["corpus_size"]) == len(docs)
    # the shards' postings are the whole corpus' postings split by document range
    assert int(s0["indptr"][-1]) + int(s1["indptr"][-1]) == int(whole["indptr"][-1])
    t = 17
    a, b = int(whole["indptr"][t]), int(whole["indptr"][t + 1])
    got = np.concatenate([s0["doc_ids"][s0["indptr"][t]: s0["indptr"][t + 1]],
                          s1["doc_ids"][s1["indptr"][t]: s1["indptr"][t + 1]] + 700])
    assert np.array_equal(got, whole["doc_ids"].numpy()[a:b])
