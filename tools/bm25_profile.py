"""BM25 kernel timing: python tools/bm25_profile.py [N_DOCS] [NQ] [key=value tunables ...]
Synthetic Zipf corpus built in HBM; times topk_dev (k = 30) over a pool of 64 queries, CUDA events."""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from rag_challenge_2_b200 import _lib, synth  # noqa: E402
from rag_challenge_2_b200.bm25 import BM25Okapi  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1_250_000
nq = int(sys.argv[2]) if len(sys.argv) > 2 else 1
for kv in sys.argv[3:]:
    key, v = kv.split("=")
    _lib.tunable_set(key, int(v))
dev = torch.device("cuda", 0)
post = synth.token_postings_torch(0, n, 50_000, dev)
idf, _ = synth.bm25_idf_from_df(post["df"].cpu().numpy(), n)
bm = BM25Okapi.from_postings(post["indptr"], post["doc_ids"], post["tf"], post["doc_len"], torch.from_numpy(idf).to(dev),
                             post["total_len"] / n, device=0, weight_bits=int(os.environ.get("B2R_BM25_BITS", "32")))
df = post["df"].cpu().numpy()
pool = synth.token_queries_np(64, vocab=50_000, seed=4243)
batches = []
for i in range(0, 64, nq):
    toks = pool[i: i + nq]
    offs = np.zeros(len(toks) + 1, np.int32)
    np.cumsum([len(t) for t in toks], out=offs[1:])
    batches.append((torch.from_numpy(np.concatenate(toks).astype(np.int32)).to(dev), torch.from_numpy(offs).to(dev),
                    sum(int(df[t]) for q in toks for t in q)))
for t, o, _ in batches[:3]:
    bm.topk_dev(t, o, 30)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
reps = 5
e0.record()
for _ in range(reps):
    for t, o, _ in batches:
        bm.topk_dev(t, o, 30)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / (reps * len(batches))
postings = sum(b[2] for b in batches) / len(batches)
print(json.dumps({"n_docs": n, "nq": nq, "ms_per_launch": round(ms, 4), "postings_per_launch": int(postings),
                  "GBps": round(postings * (4 + bm.n_postings()[1] // 8) / ms / 1e6, 1), "weight_bits": bm.n_postings()[1],
                  "tunables": sys.argv[3:]}), flush=True)
