"""Print a few fields of a bench.py JSON line read from stdin (helper for sweeps)."""
import json
import sys

label = sys.argv[1] if len(sys.argv) > 1 else ""
line = [l for l in sys.stdin.read().splitlines() if l.startswith("{")][-1]
d = json.loads(line)
print(label, "step_ms", round(d["ms_per_step"], 4), "scan_ms", round(d["roofline"]["kernel_ms"], 4), "bm25_ms",
      d["bm25"].get("kernel_ms_in_step"), "e2e_ms", round(d["e2e"]["ms_per_step"], 4), "q/s", round(d["value"], 1),
      "dense_only_ms", round(d.get("extras", {}).get("dense_single_query", {}).get("ms", 0.0), 4))
