"""Run-to-run behaviour of the host-buffer step: blocking calls vs two questions in flight, several rounds of 20 steps in one
process.  python tools/e2e_variance.py [ROWS] [tunable=value ...]"""
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import rag_challenge_2_b200 as pkg  # noqa: E402
from rag_challenge_2_b200 import _lib, hybrid, synth  # noqa: E402
from rag_challenge_2_b200.bm25 import BM25Okapi  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 10_000_000
for kv in sys.argv[2:]:
    k_, v_ = kv.split("=")
    _lib.tunable_set(k_, int(v_))
dev = torch.device("cuda", 0)
rows = synth.dense_rows_torch(0, n, 1024, dev)
post = synth.token_postings_torch(0, n, 50_000, dev)
idf, _ = synth.bm25_idf_from_df(post["df"].cpu().numpy(), n)
bm = BM25Okapi.from_postings(post["indptr"], post["doc_ids"], post["tf"], post["doc_len"], torch.from_numpy(idf).to(dev),
                             post["total_len"] / n, device=0, weight_bits=32)
index = hybrid.HybridIndex(pkg.FlatIPIndex.from_device_tensor(rows), bm, 0, world_size=1)
q = synth.dense_queries_torch(64, 1024, dev, rows).cpu().numpy()
toks = synth.token_queries_np(64, vocab=50_000, seed=4243)
steps = 20
out = {"rows": n, "tunables": sys.argv[2:], "sync_ms": [], "two_in_flight_ms": [], "per_step_ms_round0": []}
for i in range(4):
    index.search_hybrid(q[i: i + 1], [toks[i]], 30)
for rnd in range(6):
    t0 = time.perf_counter()
    for i in range(steps):
        index.search_hybrid(q[i: i + 1], [toks[i]], 30)
    out["sync_ms"].append(round((time.perf_counter() - t0) * 1e3 / steps, 4))
    t0 = time.perf_counter()
    marks = []
    pending = index.search_hybrid_submit(q[0:1], [toks[0]], 30)
    for i in range(1, steps):
        nxt = index.search_hybrid_submit(q[i: i + 1], [toks[i]], 30)
        index.search_hybrid_collect(pending)
        marks.append(time.perf_counter())
        pending = nxt
    index.search_hybrid_collect(pending)
    marks.append(time.perf_counter())
    out["two_in_flight_ms"].append(round((marks[-1] - t0) * 1e3 / steps, 4))
    if rnd == 0:
        out["per_step_ms_round0"] = [round((b - a) * 1e3, 3) for a, b in zip([t0] + marks[:-1], marks)]
print(json.dumps(out), flush=True)
