"""Multi-GPU correctness check (run under torchrun on G >= 2 GPUs): row-sharded hybrid search through
(a) the one-kernel peer-memory exchange and (b) NCCL all-gather + merge kernels must both equal the
unsharded search computed on rank 0 over the whole corpus.  Exit code 0 = identical on every rank."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import rag_challenge_2_b200 as pkg  # noqa: E402
from rag_challenge_2_b200 import hybrid, synth  # noqa: E402
from rag_challenge_2_b200.bm25 import BM25Okapi  # noqa: E402

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
n, d, V, k, nq = 16 * synth.ROW_BLOCK - 3_000, 1024, 20_000, 30, 12  # ragged last shard, none empty up to 8 ranks
bounds = hybrid.shard_bounds(n, world, 128)   # boundaries inside seeded blocks, like the bench
lo, hi = bounds[rank], bounds[rank + 1]


def build(lo, hi):
    rows = synth.dense_rows_torch(lo, hi, d, dev)
    post = synth.token_postings_torch(lo, hi, V, dev)
    return rows, post


rows, post = build(lo, hi)
df, tot = post["df"].clone(), torch.tensor([post["total_len"]], device=dev)
dist.all_reduce(df)
dist.all_reduce(tot)
idf, _ = synth.bm25_idf_from_df(df.cpu().numpy(), n)
idf_t, avgdl = torch.from_numpy(idf).to(dev), float(tot.item()) / n
flat = pkg.FlatIPIndex.from_device_tensor(rows)
bm = BM25Okapi.from_postings(post["indptr"], post["doc_ids"], post["tf"], post["doc_len"], idf_t, avgdl, device=local)
q = synth.dense_queries_torch(nq, d, dev, rows if rank == 0 else None)
dist.broadcast(q, 0)
toks = synth.token_queries_np(nq, vocab=V, seed=4243)
offs = np.zeros(nq + 1, np.int32)
np.cumsum([len(t) for t in toks], out=offs[1:])
tids = torch.from_numpy(np.concatenate(toks).astype(np.int32)).to(dev)
toffs = torch.from_numpy(offs).to(dev)

peer = hybrid.HybridIndex(flat, bm, lo, exchange="peer")
nccl = hybrid.HybridIndex(flat, bm, lo, exchange="nccl")
ok = True
for it in range(5):  # several steps: exercises the parity double-buffering and the step counter
    F1, O1 = peer.search_hybrid_dev(q, tids, toffs, k)
    F2, O2 = nccl.search_hybrid_dev(q, tids, toffs, k)
    torch.cuda.synchronize()
    ok &= bool(torch.equal(O1, O2) and torch.equal(F1, F2))
# single queries: the SIMT scan with BM25 launched beside it (programmatic dependent launch) + the exchange
t1 = torch.from_numpy(np.asarray(toks[0], np.int32)).to(dev)
o1 = torch.tensor([0, len(toks[0])], dtype=torch.int32, device=dev)
for it in range(4):
    Fa, Oa = peer.search_hybrid_dev(q[:1], t1, o1, k)
    Fb, Ob = nccl.search_hybrid_dev(q[:1], t1, o1, k)
    torch.cuda.synchronize()
    ok &= bool(torch.equal(Oa, Ob) and torch.equal(Fa, Fb) and torch.equal(Oa[0], O1[0]) and torch.equal(Fa[0], F1[0]))
# a stream of distinct single-query steps, CHAINED by programmatic dependent launch (the scan of step s + 1 queues
# behind step s, whose BM25 kernel's last CTA exchanges + merges + fuses in place): step for step the un-chained answers
per_q = []
for j in range(nq):
    tj = torch.from_numpy(np.asarray(toks[j], np.int32)).to(dev)
    oj = torch.tensor([0, len(toks[j])], dtype=torch.int32, device=dev)
    Fr, Or = nccl.search_hybrid_dev(q[j: j + 1].contiguous(), tj, oj, k)
    per_q.append((q[j: j + 1].contiguous(), tj, oj, Fr.clone(), Or.clone()))
torch.cuda.synchronize()
dist.barrier()
peer.searcher.pipeline = True
outs = []
for it in range(60):
    qj, tj, oj, _, _ = per_q[it % nq]
    Fc, Oc = peer.search_hybrid_dev(qj, tj, oj, k)
    outs.append((Fc.clone(), Oc.clone()))
torch.cuda.synchronize()
peer.searcher.peer_exchange.check()
for it, (Fc, Oc) in enumerate(outs):
    ok &= bool(torch.equal(Oc, per_q[it % nq][4]) and torch.equal(Fc, per_q[it % nq][3]))
chained_ok = ok
peer.searcher.pipeline = False
Fh, Oh = peer.search_hybrid(q.cpu().numpy(), toks, k)  # host API on every rank (SPMD)
ok &= bool(np.array_equal(Oh, O1.cpu().numpy()) and np.array_equal(Fh, F1.cpu().numpy()))
if rank == 0:
    rows_all, post_all = build(0, n)
    whole = hybrid.HybridIndex(pkg.FlatIPIndex.from_device_tensor(rows_all),
                               BM25Okapi.from_postings(post_all["indptr"], post_all["doc_ids"], post_all["tf"],
                                                       post_all["doc_len"], idf_t, avgdl, device=local), 0, world_size=1)
    F0, O0 = whole.search_hybrid_dev(q, tids, toffs, k)
    torch.cuda.synchronize()
    same = bool(torch.equal(O0, O1) and torch.equal(F0, F1))
    print("sharded == unsharded:", same, "| peer == nccl == host API:", ok, "| chained steps == un-chained:", chained_ok,
          "| launches per single-query step:", peer.searcher.launches, "| top1", O1[:4, 0].tolist())
    ok &= same
flag = torch.tensor([1 if ok else 0], device=dev)
dist.all_reduce(flag, op=dist.ReduceOp.MIN)
dist.barrier()
dist.destroy_process_group()
sys.exit(0 if flag.item() == 1 else 1)
