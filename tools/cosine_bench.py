"""Config 5: N x N cosine statistics, N x 1024 (default N = 200,000).  python tools/cosine_bench.py [N]"""
import json
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from rag_challenge_2_b200 import cosine, synth  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 200_000
d = 1024
g = torch.Generator(device="cuda").manual_seed(777)
x = torch.randn((n, d), generator=g, device="cuda") * (0.5 + 1.5 * torch.rand((n, 1), generator=g, device="cuda"))
x = x.contiguous()
cosine.cosine_stats(x[:8192].contiguous())
torch.cuda.synchronize()
times = []
for _ in range(2):          # the first full-size call pays the allocator (2 x 0.8 GB of split operands)
    t0 = time.time()
    st = cosine.cosine_stats(x)
    torch.cuda.synchronize()
    times.append(time.time() - t0)
dt = min(times)
flops = 2.0 * n * n * d / 2  # symmetry: upper triangle only
print(json.dumps({"n": n, "d": d, "seconds": round(dt, 3), "seconds_each_call": [round(t, 3) for t in times], "tflops_useful_fp32_equiv": round(flops / dt / 1e12, 1),
                  "tensor_tflops_issued_3xtf32": round(3 * flops / dt / 1e12, 1), "stats": st}))
t0 = time.time()
S, I = cosine.cosine_topk(x[: min(n, 50_000)].contiguous(), 30)
torch.cuda.synchronize()
print(json.dumps({"topk_rows": min(n, 50_000), "k": 30, "seconds": round(time.time() - t0, 3)}))
