#!/usr/bin/env bash
# Round-2 opening measurements in ONE gpurun call (1 GPU, ~4 GPU-minutes):
#   gpurun --timeout 420 -- 'bash tools/r02_first_call.sh'
# Writes everything under gpurun_out/r02/.  Nothing here runs under a profiler except the last step.
set -u
OUT=gpurun_out/r02
mkdir -p "$OUT"
# 1. the GPU suite at HEAD (parity gate before any number is read)
python -m pytest tests -x -q -m gpu > "$OUT/pytest_gpu.log" 2>&1; echo "pytest rc=$?" | tee -a "$OUT/pytest_gpu.log"
# 2. headline line
python bench.py --steps 100 --warmup 5 > "$OUT/bench_n1.json" 2> "$OUT/bench_n1.err"
# 3. the 8-GPU shard on one GPU (1.25M rows): where do the ~76 us above the scan go?  
for variant in "" "--no-overlap" "--tunable bm25.v1=2" "--no-bm25-events"; do
  python bench.py --rows 1261568 --steps 300 --warmup 10 --no-cpu-baseline $variant \
    | python -c "import json,sys; d=json.loads(sys.stdin.readline()); print(json.dumps({'variant': '$variant', 'ms_per_step': d['ms_per_step'], 'e2e_ms': d['e2e']['ms_per_step'], 'scan_ms': d['roofline']['kernel_ms'], 'dense_single_ms': d['extras']['dense_single_query']['ms']}))" \
    >> "$OUT/shard_breakdown.jsonl" 2>> "$OUT/shard_breakdown.err"
done
# 4. scan alone at shard size for ring depth / tile size (fixed costs: ring fill, tail)
for t in "flat.stages=3" "flat.stages=6" "flat.tile_rows=16" "flat.tile_rows=64"; do
  echo "$t $(python - "$t" <<'PY'
import sys
sys.path.insert(0, ".")
from rag_challenge_2_b200 import _lib
k, v = sys.argv[1].split("=")
_lib.tunable_set(k, int(v))
sys.argv = ["flat_profile.py", "1", "1261568", "50"]
__file__ = "tools/flat_profile.py"
exec(open(__file__).read())
PY
)" >> "$OUT/flat_shard_tunables.txt" 2>&1
done
# 5. config 2 and BM25 batches (host path after the direct record build; K4s at 64 queries)
python tools/yeardb_bench.py --reps 3 > "$OUT/yeardb.json" 2>&1
python tools/bm25_profile.py 10000000 64 > "$OUT/bm25_64.txt" 2>&1
# 6. launch list of the bench command (profiles/r02_launches_bench_n1.csv)
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file "$OUT/launches_bench_n1.csv" \
  python bench.py --steps 2 --warmup 1 --no-cpu-baseline > "$OUT/ncu_bench.log" 2>&1
ls -la "$OUT"
