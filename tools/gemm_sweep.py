"""K2 timing sweep: python tools/gemm_sweep.py [N_ROWS] [nq,nq,...] [key=value tunables ...]
Whole search (prep + scans + selects + rescoring + redo check), CUDA events on torch's stream."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import rag_challenge_2_b200 as pkg  # noqa: E402
from rag_challenge_2_b200 import _lib, synth  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 4_000_000
nqs = [int(v) for v in sys.argv[2].split(",")] if len(sys.argv) > 2 else [8, 16, 32, 64, 128, 256, 1024]
use_bf16 = False
for kv in sys.argv[3:]:
    key, v = kv.split("=")
    if key == "bf16":
        use_bf16 = bool(int(v))
    else:
        _lib.tunable_set(key, int(v))
rows = synth.dense_rows_torch(0, n, 1024, "cuda")
index = pkg.FlatIPIndex.from_device_tensor(rows)
if use_bf16:
    index.enable_bf16_candidates()
out = []
for nq in nqs:
    q = synth.dense_queries_torch(nq, 1024, "cuda", rows)
    D = torch.empty((nq, 30), dtype=torch.float32, device="cuda")
    I = torch.empty((nq, 30), dtype=torch.int64, device="cuda")
    for _ in range(2):
        index.search_dev(q, 30, D, I)
    torch.cuda.synchronize()
    iters = 5 if nq <= 256 else 2
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        index.search_dev(q, 30, D, I)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / iters
    passes = -(-nq // 128)
    rec = {"nq": nq, "n": n, "ms": round(ms, 4), "qps": round(nq / ms * 1e3, 1),
           "GBps_per_pass": round(passes * n * 4096 / ms / 1e6, 1), "TFLOPs": round(2 * n * 1024 * nq / ms / 1e9, 1),
           "launches": index.last_launch_count()}
    out.append(rec)
    print(json.dumps(rec), flush=True)
os.makedirs("gpurun_out", exist_ok=True)
with open("gpurun_out/gemm_sweep.json", "a") as f:
    f.write(json.dumps({"argv": sys.argv[1:], "results": out}) + "\n")
