"""How much do the tensor-core batch searches move with the power cap?  Runs 256- and 1024-query batches over the 10M-row
index back to back for a few seconds each and reports per-iteration times next to the SM clock / power samples taken
meanwhile (BENCH and SCALE disagreed by 14 % on the same 256-query batch).
    python tools/power_characterise.py [ROWS] [SECONDS]"""
import json
import os
import subprocess
import sys
import threading
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import rag_challenge_2_b200 as pkg  # noqa: E402
from rag_challenge_2_b200 import synth  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 10_000_000
seconds = float(sys.argv[2]) if len(sys.argv) > 2 else 3.0
rows = synth.dense_rows_torch(0, n, 1024, "cuda")
index = pkg.FlatIPIndex.from_device_tensor(rows)
samples = []
stop = False


def smi():
    p = subprocess.Popen(["nvidia-smi", "--id=0", "--query-gpu=clocks.sm,power.draw,clocks_event_reasons.sw_power_cap",
                          "--format=csv,noheader,nounits", "-lms", "100"], stdout=subprocess.PIPE, text=True)
    for line in p.stdout:
        if stop:
            break
        parts = [x.strip() for x in line.split(",")]
        try:
            samples.append((time.time(), float(parts[0]), float(parts[1]), parts[2]))
        except (ValueError, IndexError):
            pass
    p.terminate()


threading.Thread(target=smi, daemon=True).start()
out = {"rows": n, "seconds_per_case": seconds, "cases": {}}
for name, nq, bf16 in (("1 query (scan, HBM-bound)", 1, False), ("256 queries tf32", 256, False), ("1024 queries tf32", 1024, False),
                       ("1024 queries bf16 copy", 1024, True)):
    q = synth.dense_queries_torch(nq, 1024, "cuda", rows, seed=600 + nq)
    index.enable_bf16_candidates(bf16)
    for _ in range(2):
        index.search_dev(q, 30)
    torch.cuda.synchronize()
    time.sleep(1.0)                      # let the clocks recover: every case starts from an idle GPU
    t_case0 = time.time()
    times = []
    while time.time() - t_case0 < seconds:
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        index.search_dev(q, 30)
        b.record()
        torch.cuda.synchronize()
        times.append(a.elapsed_time(b))
    t_case1 = time.time()
    mine = [s for s in samples if t_case0 <= s[0] <= t_case1]
    k = max(1, len(times) // 10)
    out["cases"][name] = {
        "iterations": len(times), "ms_first_10pct": sum(times[:k]) / k, "ms_last_10pct": sum(times[-k:]) / k,
        "ms_min": min(times), "ms_max": max(times), "drift_pct": 100.0 * (sum(times[-k:]) / sum(times[:k]) - 1.0),
        "sm_mhz_first_last": [mine[0][1], mine[-1][1]] if mine else None,
        "power_w_first_last": [mine[0][2], mine[-1][2]] if mine else None,
        "sw_power_cap_samples": sum(1 for s in mine if s[3].lower().startswith("active")), "smi_samples": len(mine)}
index.enable_bf16_candidates(False)
stop = True
print(json.dumps(out), flush=True)
