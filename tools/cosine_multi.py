"""Config 5 on several GPUs (run under torchrun): N x N cosine statistics with the COLUMNS sharded over the ranks.

    python -m torch.distributed.run --nproc-per-node G tools/cosine_multi.py [N]

Every rank holds X (0.82 GB at N = 200k), reduces the pairs (i, j), i < j, whose column j lies in its range, and the
accumulators (sum, sum of squares, count, histogram; min / max) are all-reduced — SURVEY 8e(iii).  Work grows with j
(the pairs below the diagonal), so the column ranges have equal AREA: boundary r at n * sqrt(r / G), on 128-column tiles.
Rank 0 also runs the whole matrix on its own GPU and checks that the sharded statistics equal it.  Prints one JSON line."""
import json
import math
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from rag_challenge_2_b200 import cosine  # noqa: E402

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
n = int(sys.argv[1]) if len(sys.argv) > 1 else 200_000
d = 1024
g = torch.Generator(device=dev).manual_seed(777)   # the same X on every rank
x = (torch.randn((n, d), generator=g, device=dev) * (0.5 + 1.5 * torch.rand((n, 1), generator=g, device=dev))).contiguous()
bounds = [min(n, int(round(n * math.sqrt(r / world) / 128.0)) * 128) for r in range(world)] + [n]
lo, hi = bounds[rank], bounds[rank + 1]
cosine.cosine_stats(x[:8192].contiguous())          # warm-up (allocations, tensor maps)
torch.cuda.synchronize()
dist.barrier()


def sharded():
    stats, hist = cosine.cosine_stats(x, lo, hi, finalize=False)
    packed = torch.stack([stats[0], stats[1], stats[4]])
    dist.all_reduce(packed)
    dist.all_reduce(hist)
    mn, mx = stats[2].clone(), stats[3].clone()
    dist.all_reduce(mn, op=dist.ReduceOp.MIN)
    dist.all_reduce(mx, op=dist.ReduceOp.MAX)
    return cosine.finalize_stats(np.array([float(packed[0]), float(packed[1]), float(mn), float(mx), float(packed[2]), 0.0]),
                                 hist.cpu().numpy())


times = []
for _ in range(3):
    torch.cuda.synchronize()
    dist.barrier()
    t0 = time.perf_counter()
    got = sharded()
    torch.cuda.synchronize()
    times.append(time.perf_counter() - t0)
t = torch.tensor([min(times)], dtype=torch.float64, device=dev)
dist.all_reduce(t, op=dist.ReduceOp.MAX)
ok = True
if rank == 0:
    t0 = time.perf_counter()
    whole = cosine.cosine_stats(x)
    torch.cuda.synchronize()
    t_one = time.perf_counter() - t0
    t0 = time.perf_counter()
    whole = cosine.cosine_stats(x)
    torch.cuda.synchronize()
    t_one = min(t_one, time.perf_counter() - t0)
    ok = got["count"] == whole["count"] and got["max_similarity"] == whole["max_similarity"] and \
        got["min_similarity"] == whole["min_similarity"] and got["median_similarity"] == whole["median_similarity"]
    for key in ("avg_similarity", "std_similarity"):      # float64 sums of fp32 chunk sums, added in another order
        ok = ok and abs(got[key] - whole[key]) <= 1e-9 * abs(whole[key]) + 1e-12
    useful = 2.0 * n * n * d / 2
    print(json.dumps({"config": "N x N cosine statistics, columns sharded over the ranks (equal-area ranges), stats all-reduced",
                      "n": n, "d": d, "n_gpus": world, "seconds": float(t.item()), "seconds_1gpu_same_box": t_one,
                      "speedup": t_one / float(t.item()), "tf32_tflops_issued_all_gpus": 3 * useful / float(t.item()) / 1e12,
                      "equals_single_gpu": bool(ok), "column_bounds": bounds, "stats": got}), flush=True)
flag = torch.tensor([1 if ok else 0], device=dev)
dist.all_reduce(flag, op=dist.ReduceOp.MIN)
dist.barrier()
dist.destroy_process_group()
sys.exit(0 if flag.item() == 1 else 1)
