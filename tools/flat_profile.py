"""Tiny harness for ncu: a few searches over a resident index.  python tools/flat_profile.py NQ [N] [ITERS] [chain]
chain: launch every search with B2R_LAUNCH_OVERLAP_PREVIOUS (each scan may start while the previous one finishes)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import rag_challenge_2_b200 as pkg  # noqa: E402
from rag_challenge_2_b200 import synth  # noqa: E402

nq = int(sys.argv[1]) if len(sys.argv) > 1 else 1
n = int(sys.argv[2]) if len(sys.argv) > 2 else 4_000_000
iters = int(sys.argv[3]) if len(sys.argv) > 3 else 4
chain = len(sys.argv) > 4 and sys.argv[4] == "chain"
rows = synth.dense_rows_torch(0, n, 1024, "cuda")
index = pkg.FlatIPIndex.from_device_tensor(rows)
q = synth.dense_queries_torch(nq, 1024, "cuda", rows)
D = torch.empty((nq, 30), dtype=torch.float32, device="cuda")
I = torch.empty((nq, 30), dtype=torch.int64, device="cuda")
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
index.search_dev(q, 30, D, I)
torch.cuda.synchronize()
e0.record()
for _ in range(iters):
    index.search_dev(q, 30, D, I, overlap_previous=chain)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / iters
print(f"{'chained ' if chain else ''}nq={nq} n={n} ms={ms:.4f} GB/s={n * 4096 / ms / 1e6:.1f} top1={I[:, 0].tolist()}")
