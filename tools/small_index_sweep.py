"""Latency of index.search on report-sized indexes (1e3..1e5 chunks), SIMT passes vs the tensor-core path.
python tools/small_index_sweep.py  ->  one JSON line per (n, nq): host-API wall time per call."""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import rag_challenge_2_b200 as pkg  # noqa: E402
from rag_challenge_2_b200 import _lib  # noqa: E402

rng = np.random.default_rng(0)
rows_out = []
for n in (2000, 10000, 50000, 200000):
    xb = rng.standard_normal((n, 1024), dtype=np.float32)
    xb /= np.linalg.norm(xb, axis=1, keepdims=True)
    index = pkg.FlatIPIndex(1024)
    index.add(xb)
    for nq in (1, 2, 3, 4, 5, 8, 12, 16, 32):
        xq = xb[:nq] + 0.01 * rng.standard_normal((nq, 1024), dtype=np.float32)
        rec = {"n": n, "nq": nq}
        for name, flag in (("simt_us", 1), ("tensor_us", 0)):
            _lib.tunable_set("flat.no_gemm", flag)
            for _ in range(3):
                index.search(xq, 30)
            t0 = time.perf_counter()
            for _ in range(20):
                index.search(xq, 30)
            rec[name] = round((time.perf_counter() - t0) / 20 * 1e6, 1)
        _lib.tunable_set("flat.no_gemm", 0)
        rows_out.append(rec)
        print(json.dumps(rec), flush=True)
os.makedirs("gpurun_out", exist_ok=True)
json.dump(rows_out, open("gpurun_out/small_index_sweep.json", "w"))
