for dyn in -1 4 1 2 8 16; do
python - $dyn <<'PY'
import sys, os
sys.path.insert(0, ".")
import torch
import rag_challenge_2_b200 as pkg
from rag_challenge_2_b200 import _lib, synth
dyn = int(sys.argv[1])
_lib.tunable_set("flat.dynamic_tiles", dyn)
for n in (1250048, 10000000):
    rows = synth.dense_rows_torch(0, n, 1024, "cuda")
    index = pkg.FlatIPIndex.from_device_tensor(rows)
    q = synth.dense_queries_torch(4, 1024, "cuda", rows)
    D = torch.empty((1, 30), dtype=torch.float32, device="cuda"); I = torch.empty((1, 30), dtype=torch.int64, device="cuda")
    ref = None
    for qi in range(4):
        index.search_dev(q[qi:qi+1].contiguous(), 30, D, I)
    torch.cuda.synchronize()
    iters = 300 if n < 5_000_000 else 60
    best = 1e9
    for rep in range(3):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(iters):
            index.search_dev(q[:1].contiguous(), 30, D, I)
        e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1) / iters)
    print(f"dyn={dyn} n={n} ms={best:.4f} GB/s={n*4096/best/1e6:.1f} top1={I[0,0].item()} d0={D[0,0].item():.6f}")
    del index, rows
PY
done
