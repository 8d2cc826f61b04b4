"""GPU, BASELINE's full size (10M x 1024 fp32 rows + BM25 over 10M chunks, vocab 50k — the bench corpus): parity
through size-independent properties, since the CPU oracle cannot touch 41 GB in test time.
  dense  — every stored row finds itself first (unit rows); scores are linear in the query; returned scores equal a
           float64 recomputation of those rows; no row outside the result beats the k-th score (cuBLAS fp32 scores of
           all rows — checker only); the tensor-core batch path returns the single-query path's answers; three row
           shards merged == the whole index, bit for bit.
  BM25   — the default (float64-weight) index returns rank_bm25's un-rounded float64 scores BIT FOR BIT and the stable-sort
           top-30 ids, checked against a float64 torch restatement of the formula over the same postings; the float32-weight
           index (the bench layout) is measured against the same un-rounded scores: max relative deviation <= 1e-6, ids
           equal except swaps inside that band (the report goes to gpurun_out/full_size_bm25_parity.json).
  hybrid — the one-launch-sequence step equals fusing the separately computed parts, and the host-buffer call equals it.
B2R_FULL_ROWS overrides the size (debugging); skipped when the GPU has less than 120 GB free."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

N = int(os.environ.get("B2R_FULL_ROWS", "10000000"))
D, V, K = 1024, 50_000, 30


@pytest.fixture(scope="module")
def corpus(b2r):
    import torch

    from rag_challenge_2_b200 import synth
    from rag_challenge_2_b200.bm25 import BM25Okapi

    free, _ = torch.cuda.mem_get_info(0)
    if N >= 5_000_000 and free < 120 * 2**30:
        pytest.skip(f"needs 120 GB of free HBM, {free / 2**30:.0f} GB available")
    dev = torch.device("cuda", 0)
    rows = synth.dense_rows_torch(0, N, D, dev)
    flat = b2r.FlatIPIndex.from_device_tensor(rows)
    post = synth.token_postings_torch(0, N, V, dev)
    idf, _ = synth.bm25_idf_from_df(post["df"].cpu().numpy(), N)
    avgdl = post["total_len"] / N
    bm25 = BM25Okapi.from_postings(post["indptr"], post["doc_ids"], post["tf"], post["doc_len"],
                                   torch.from_numpy(idf).to(dev), avgdl, device=0)
    yield {"rows": rows, "flat": flat, "post": post, "idf": idf, "avgdl": avgdl, "bm25": bm25, "dev": dev}
    bm25.close()
    flat.close()
    del rows, post
    torch.cuda.empty_cache()


def test_dense_properties_full_size(b2r, corpus):
    import torch

    from rag_challenge_2_b200 import hybrid, synth

    rows, flat, dev = corpus["rows"], corpus["flat"], corpus["dev"]
    # K1 (two queries per pass): self match, linearity
    pick = torch.tensor([0, N - 1], device=dev)
    q = rows[pick].contiguous()
    D1, I1 = flat.search_dev(q, K)
    D2, I2 = flat.search_dev((2.0 * q).contiguous(), K)
    torch.cuda.synchronize()
    assert torch.equal(I1[:, 0], pick) and torch.allclose(D1[:, 0], torch.ones(2, device=dev), atol=1e-5)
    assert torch.equal(I1, I2) and torch.allclose(D2, 2.0 * D1, rtol=1e-6)
    # single queries (the headline path) on mixed queries: exact scores, k-th score bound, order
    qs = synth.dense_queries_torch(6, D, dev, rows)
    singles = [flat.search_dev(qs[i: i + 1].contiguous(), K) for i in range(6)]
    Ds = torch.cat([s[0] for s in singles])
    Is = torch.cat([s[1] for s in singles])
    assert int(Is.min()) >= 0 and int(Is.max()) < N
    ref = (rows[Is.reshape(-1)].double() * qs.double().repeat_interleave(K, dim=0)).sum(1).reshape(6, K)
    assert torch.allclose(Ds.double(), ref, rtol=1e-5, atol=1e-6)
    assert torch.all(Ds[:, 1:] <= Ds[:, :-1])
    full = (rows @ qs.T).T                                     # [6, N] fp32 by cuBLAS: checker only
    assert torch.all((full > Ds[:, -1:] + 1e-5).sum(1) <= K - 1)
    top_ref = torch.topk(full, K, dim=1)
    clear = (top_ref.values[:, :-1] - top_ref.values[:, 1:]).min(dim=1).values > 1e-5   # no near-ties: ids must match
    assert torch.equal(Is[clear], top_ref.indices[clear])
    del full
    # K2 (tensor cores, 64 queries per pass) returns the single-query answers
    q64 = synth.dense_queries_torch(64, D, dev, rows, seed=7)
    q64[:6] = qs
    Db, Ib = flat.search_dev(q64.contiguous(), K)
    torch.cuda.synchronize()
    assert torch.equal(Ib[:6], Is)
    assert torch.allclose(Db[:6], Ds, rtol=2e-6, atol=1e-7)
    # three row shards (zero-copy views) merged == the whole index
    bounds = [0, (N // 3 // 16) * 16, (2 * N // 3 // 16) * 16 + 16, N]
    parts = []
    for lo, hi in zip(bounds[:-1], bounds[1:]):
        shard = b2r.FlatIPIndex.from_device_tensor(rows[lo:hi])
        parts.append(shard.search_dev(qs[:2].contiguous(), K, id_offset=lo))
        torch.cuda.synchronize()
        shard.close()
    Dm, Im = hybrid.merge_topk(torch.stack([p[0] for p in parts]), torch.stack([p[1] for p in parts]), K)
    Dw, Iw = flat.search_dev(qs[:2].contiguous(), K)
    torch.cuda.synchronize()
    assert torch.equal(Im, Iw) and torch.equal(Dm, Dw)


@pytest.mark.parametrize("nq", [1024, 4096])
def test_large_batches_full_size(b2r, corpus, nq):
    """BASELINE config 3's upper batch sizes (query batch 1-4096) over the full 10M rows: the tensor-core batch path (256
    queries per pass on CTA pairs; tf32 candidates, then the bf16-copy candidates) must return the single-query scan's answers
    — ids identical, scores to the last ulps (both rescore in fp32) — and every returned score must equal a float64
    recomputation of that row; a sample of queries is also checked against cuBLAS scores of ALL rows (checker only)."""
    import torch

    from rag_challenge_2_b200 import synth

    rows, flat, dev = corpus["rows"], corpus["flat"], corpus["dev"]
    q = synth.dense_queries_torch(nq, D, dev, rows, seed=11 + nq)
    sample = torch.linspace(0, nq - 1, 24, device=dev).long()
    singles = [flat.search_dev(q[i: i + 1].contiguous(), K) for i in sample.tolist()]
    Ds = torch.cat([s[0] for s in singles])
    Is = torch.cat([s[1] for s in singles])
    full = (rows @ q[sample].T).T                                   # [24, N] fp32 by cuBLAS: checker only
    top_ref = torch.topk(full, K, dim=1)
    kth = top_ref.values[:, -1:].clone()
    clear = (top_ref.values[:, :-1] - top_ref.values[:, 1:]).min(dim=1).values > 1e-5
    del full
    for bf16 in (False, True):
        flat.enable_bf16_candidates(bf16)
        try:
            Db, Ib = flat.search_dev(q, K)
            torch.cuda.synchronize()
        finally:
            flat.enable_bf16_candidates(False)
        assert int(Ib.min()) >= 0 and int(Ib.max()) < N
        assert torch.all(Db[:, 1:] <= Db[:, :-1])
        for lo in range(0, nq, 512):                                # float64 recomputation of every returned score
            hi = min(nq, lo + 512)
            got = rows[Ib[lo:hi].reshape(-1)].double()
            ref = (got * q[lo:hi].double().repeat_interleave(K, dim=0)).sum(1).reshape(hi - lo, K)
            assert torch.allclose(Db[lo:hi].double(), ref, rtol=1e-5, atol=1e-6), (bf16, lo)
        assert torch.equal(Ib[sample], Is), ("bf16" if bf16 else "tf32")
        assert torch.allclose(Db[sample], Ds, rtol=2e-6, atol=1e-7)
        assert torch.all(Db[sample][:, -1:] >= kth - 1e-5)          # nothing outside the result beats the k-th score
        assert torch.equal(Ib[sample][clear], top_ref.indices[clear])
        # every query that is a noisy copy of a stored row (even positions: SURVEY 8d) must rank a score > 0.9 first
        assert float(Db[0::2, 0].min()) > 0.9


def _bm25_reference_scores(corpus, terms, k1=1.5, b=0.75):
    """rank_bm25's get_scores (src/retrieval.py:262) over the CSR postings in UN-ROUNDED float64, numpy's operation order
    (torch's elementwise float64 ops are the same IEEE operations): nothing here knows how the index stores its weights."""
    import torch

    post, idf, avgdl = corpus["post"], corpus["idf"], corpus["avgdl"]
    score = torch.zeros(N, dtype=torch.float64, device=corpus["dev"])
    dl = post["doc_len"]
    avgdl = torch.tensor(avgdl, dtype=torch.float64, device=corpus["dev"])   # a tensor: torch turns division by a Python
    for t in terms:                                                           # scalar into a multiplication by 1/x
        lo, hi = int(post["indptr"][t]), int(post["indptr"][t + 1])
        d = post["doc_ids"][lo:hi].long()
        f = post["tf"][lo:hi].double()
        w = f * (k1 + 1) / (f + k1 * (1 - b + b * dl[d].double() / avgdl))
        score[d] += float(idf[t]) * w                              # documents are unique within a term
    return score


def _stable_topk(ref, k):
    """Python's stable descending sort (src/retrieval.py:265): score descending, equal scores by ascending id."""
    import torch

    kth = torch.topk(ref, k).values[-1]
    cand = torch.nonzero(ref >= kth).reshape(-1)                   # ascending ids
    order = torch.sort(ref[cand], descending=True, stable=True).indices[:k]
    return ref[cand][order], cand[order]


def test_bm25_and_hybrid_full_size(b2r, corpus):
    import json

    import torch

    from rag_challenge_2_b200 import hybrid, synth
    from rag_challenge_2_b200.bm25 import BM25Okapi

    bm25, flat, rows, dev = corpus["bm25"], corpus["flat"], corpus["rows"], corpus["dev"]
    assert bm25.n_postings()[1] == 64                              # the default layout: rank_bm25's float64 weights
    post = corpus["post"]
    bm32 = BM25Okapi.from_postings(post["indptr"], post["doc_ids"], post["tf"], post["doc_len"],
                                   torch.from_numpy(corpus["idf"]).to(dev), corpus["avgdl"], device=0, weight_bits=32)
    queries = synth.token_queries_np(16, vocab=V, seed=4243)
    offs = np.zeros(len(queries) + 1, np.int32)
    np.cumsum([len(t) for t in queries], out=offs[1:])
    tids = torch.from_numpy(np.concatenate(queries).astype(np.int32)).to(dev)
    toffs = torch.from_numpy(offs).to(dev)
    S, J = bm25.topk_dev(tids, toffs, K)
    S32, J32 = bm32.topk_dev(tids, toffs, K)
    torch.cuda.synchronize()
    report = {"n_docs": N, "vocab": V, "k": K, "queries": len(queries), "f64_scores_bit_identical": True,
              "f64_ids_identical": True, "f32_max_rel_dev": 0.0, "f32_ids_identical_queries": 0, "f32_id_swaps_in_band": 0}
    for qi, terms in enumerate(queries):
        ref = _bm25_reference_scores(corpus, terms)
        vals, ids = _stable_topk(ref, K)
        # float64-weight index: the reference's float64 arithmetic bit for bit, ids and tie order included
        assert torch.equal(S[qi], vals), (qi, (S[qi] - vals).abs().max())
        assert torch.equal(J[qi], ids), qi
        # float32-weight index (the bench's 8 B / posting layout) against the SAME un-rounded float64 scores
        rel = ((S32[qi] - ref[J32[qi]]).abs() / ref[J32[qi]].abs().clamp_min(1e-300)).max().item()
        report["f32_max_rel_dev"] = max(report["f32_max_rel_dev"], rel)
        assert rel <= 1e-6, (qi, rel)
        if torch.equal(J32[qi], ids):
            report["f32_ids_identical_queries"] += 1
        else:                                                      # only swaps inside the 1e-6 band around equal scores
            diff = set(J32[qi].tolist()) ^ set(ids.tolist())
            band = 2e-6 * abs(float(vals[-1]))
            assert all(abs(float(ref[j]) - float(vals[-1])) <= band for j in diff), (qi, sorted(diff))
            pos = torch.nonzero(J32[qi] != ids).reshape(-1)
            assert all(abs(float(ref[J32[qi, p]]) - float(vals[p])) <= 2e-6 * abs(float(vals[p])) for p in pos.tolist())
            report["f32_id_swaps_in_band"] += int(len(pos))
    try:
        os.makedirs("gpurun_out", exist_ok=True)
        with open("gpurun_out/full_size_bm25_parity.json", "w") as f:
            json.dump(report, f)
    except OSError:
        pass
    print("full-size BM25 parity:", report)
    bm32.close()
    # hybrid: the fused step == fusing the parts; the host-buffer call == the resident one
    index = hybrid.HybridIndex(flat, bm25, 0, world_size=1)
    q = synth.dense_queries_torch(4, D, dev, rows, seed=5)
    for qi in range(4):
        q1 = q[qi: qi + 1].contiguous()
        t1 = tids[offs[qi]: offs[qi + 1]].contiguous()
        o1 = torch.tensor([0, len(queries[qi])], dtype=torch.int32, device=dev)
        F, O = index.search_hybrid_dev(q1, t1, o1, K)
        Dq, Iq = flat.search_dev(q1, K)
        Sq, Jq = bm25.topk_dev(t1, o1, K)
        Fr, Or = hybrid.hybrid_fuse(Dq, Iq, Sq, Jq, K)
        torch.cuda.synchronize()
        assert torch.equal(O, Or) and torch.equal(F, Fr)
        assert torch.equal(Jq[0], J[qi]) and torch.equal(Sq[0], S[qi])      # one query == its row of the batch
        Fh, Oh = index.search_hybrid(q1.cpu().numpy(), [queries[qi]], K)
        assert np.array_equal(Oh, O.cpu().numpy()) and np.array_equal(Fh, F.cpu().numpy())
    index.close()
