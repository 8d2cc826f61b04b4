#!/usr/bin/env bash
# Round-2 measurements of the fixed costs that limit the 8-GPU step, on ONE GPU holding an 8-GPU shard (1,250,048 rows):
#   gpurun --timeout 900 -- 'bash tools/r02_shard.sh'
# Writes gpurun_out/r02_shard/*.  Nothing here runs under a profiler.
set -u
OUT=gpurun_out/r02_shard
mkdir -p "$OUT"
timeout 600 python -m pytest tests/test_bm25_gpu.py tests/test_hybrid_gpu.py -x -q 2>&1 | tail -5
# BM25 alone (streaming kernel, single-barrier empty check)
for args in "10000000 1" "10000000 64" "1250048 1" "10000000 64 bm25.ctas_per_sm=1"; do
  timeout 200 python tools/bm25_profile.py $args | tee -a "$OUT/bm25.jsonl"
done
# K1 alone: fixed cost = time at a size whose streaming time is negligible
for n in 9472 151552 1250048 10000000; do
  timeout 200 python tools/flat_profile.py 1 $n 200 | tee -a "$OUT/flat_fixed_cost.txt"
done
# the hybrid step at shard size: where do the microseconds above the scan go?
for variant in "" "--no-overlap" "--tunable bm25.v1=2" "--tunable bm25.v1=1"; do
  timeout 300 python bench.py --rows 1250048 --steps 300 --warmup 10 --no-cpu-baseline --no-extras $variant \
    | python -c "import json,sys; d=json.loads([l for l in sys.stdin if l.startswith('{')][-1]); print(json.dumps({'variant': '$variant', 'ms_per_step': d['ms_per_step'], 'e2e_ms': d['e2e']['ms_per_step'], 'scan_ms': d['roofline']['kernel_ms'], 'dense_single_ms': d['extras']['dense_single_query']['ms']}))" \
    | tee -a "$OUT/shard_breakdown.jsonl"
done
