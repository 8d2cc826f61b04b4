"""BM25 index build on the GPU: the library's build kernels vs the torch sort formulation.  python tools/build_bench.py [N_DOCS]"""
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from rag_challenge_2_b200 import synth  # noqa: E402
from rag_challenge_2_b200.bm25 import token_stream_statistics  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 2_000_000
V = 50_000
dev = torch.device("cuda", 0)
g = torch.Generator(device=dev).manual_seed(5)
z = torch.randn((n,), generator=g, device=dev, dtype=torch.float64)
lens = torch.clamp(torch.round(torch.exp(np.log(120.0) + 0.5 * z)), 8, 512).to(torch.int64)
offs = torch.zeros(n + 1, dtype=torch.int64, device=dev)
offs[1:] = torch.cumsum(lens, 0)
total = int(offs[-1])
cdf = torch.from_numpy(synth.zipf_cdf(V)).to(dev)
tok = torch.empty(total, dtype=torch.int32, device=dev)
for lo in range(0, total, 1 << 26):
    hi = min(total, lo + (1 << 26))
    u = torch.rand((hi - lo,), generator=g, device=dev, dtype=torch.float64)
    tok[lo:hi] = torch.clamp(torch.searchsorted(cdf, u), max=V - 1).to(torch.int32)
out = {"n_docs": n, "tokens": total, "vocab": V}
for name, use in (("kernels", True), ("torch_sort", False)):
    torch.cuda.synchronize()
    best = 1e30
    for _ in range(2):
        t0 = time.perf_counter()
        st = token_stream_statistics(tok, offs, V, device=dev, use_kernels=use)
        torch.cuda.synchronize()
        best = min(best, time.perf_counter() - t0)
    out[name + "_s"] = round(best, 3)
    out[name + "_postings"] = int(st["doc_ids"].shape[0])
    if use:
        ref = st
    else:
        out["identical"] = bool(torch.equal(ref["doc_ids"], st["doc_ids"]) and torch.equal(ref["tf"], st["tf"]) and
                                torch.equal(ref["indptr"], st["indptr"]) and np.array_equal(ref["idf"], st["idf"]))
    del st
print(json.dumps(out), flush=True)
