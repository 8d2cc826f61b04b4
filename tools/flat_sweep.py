"""Sweep the K1 scan tunables on one GPU and print achieved GB/s per configuration.
Usage (under gpurun): python tools/flat_sweep.py [n_rows] [d]"""
import itertools
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import rag_challenge_2_b200 as pkg  # noqa: E402
from rag_challenge_2_b200 import _lib  # noqa: E402


def time_search(index, q, k, iters=5, warm=2):
    D = torch.empty((q.shape[0], k), dtype=torch.float32, device="cuda")
    I = torch.empty((q.shape[0], k), dtype=torch.int64, device="cuda")
    for _ in range(warm):
        index.search_dev(q, k, D, I)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        index.search_dev(q, k, D, I)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 4_000_000
    d = int(sys.argv[2]) if len(sys.argv) > 2 else 1024
    g = torch.Generator(device="cuda").manual_seed(1234)
    rows = torch.empty((n, d), dtype=torch.float32, device="cuda")
    step = 1 << 20
    for lo in range(0, n, step):
        hi = min(n, lo + step)
        blk = torch.randn((hi - lo, d), generator=g, device="cuda")
        rows[lo:hi] = blk / blk.norm(dim=1, keepdim=True)
    index = pkg.FlatIPIndex.from_device_tensor(rows)
    gq = torch.Generator(device="cuda").manual_seed(99)
    qall = torch.randn((8, d), generator=gq, device="cuda")
    qall /= qall.norm(dim=1, keepdim=True)
    gbytes = n * d * 4 / 1e9
    for kv in sys.argv[4:]:   # key=value tunables for the "defaults" pass
        key, v = kv.split("=")
        _lib.tunable_set(key, int(v))
    for nq in (1, 2, 3, 4, 8):
        ms = time_search(index, qall[:nq].contiguous(), 30)
        print(json.dumps({"defaults": 1, "tunables": sys.argv[4:], "nq": nq, "ms": round(ms, 4), "gbps": round(gbytes / ms * 1e3, 1)}), flush=True)
    if len(sys.argv) > 3 and sys.argv[3] == "defaults":
        return
    results = []
    grid = {
        "flat.consumer_warps": [4, 8],
        "flat.ctas_per_sm": [1, 2],
        "flat.stages": [2, 3, 4, 6],
        "flat.tile_rows": [4, 8, 16],
    }
    keys = list(grid)
    for combo in itertools.product(*[grid[k] for k in keys]):
        cfg = dict(zip(keys, combo))
        for kk, v in cfg.items():
            _lib.tunable_set(kk, v)
        for nq in (1, 4):
            try:
                ms = time_search(index, qall[:nq].contiguous(), 30)
            except Exception as e:  # noqa: BLE001
                print("skip", cfg, nq, e)
                continue
            rec = dict(cfg, nq=nq, ms=round(ms, 4), gbps=round(gbytes / ms * 1e3, 1))
            results.append(rec)
            print(json.dumps(rec), flush=True)
    for kk in keys:
        _lib.tunable_set(kk, 0)
    _lib.tunable_set("flat.force_generic", 1)
    for nq in (1, 4):
        ms = time_search(index, qall[:nq].contiguous(), 30)
        print(json.dumps({"generic": 1, "nq": nq, "ms": round(ms, 4), "gbps": round(gbytes / ms * 1e3, 1)}), flush=True)
    _lib.tunable_set("flat.force_generic", 0)
    best = {}
    for r in results:
        b = best.get(r["nq"])
        if b is None or r["ms"] < b["ms"]:
            best[r["nq"]] = r
    print("BEST", json.dumps(best))
    os.makedirs("gpurun_out", exist_ok=True)
    with open("gpurun_out/flat_sweep.json", "w") as f:
        json.dump({"n": n, "d": d, "results": results, "best": best}, f, indent=1)


if __name__ == "__main__":
    main()
