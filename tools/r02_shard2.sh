#!/usr/bin/env bash
# Round-2, second shard call: chained steps (programmatic dependent launch) and the K5 drain/statistics split.
set -u
OUT=gpurun_out/r02_shard
mkdir -p "$OUT"
timeout -k 10 600 python -m pytest tests/test_hybrid_gpu.py tests/test_cosine_gpu.py tests/test_flat_ip_gpu.py -x -q 2>&1 | tail -5
for variant in "" "--no-pipeline"; do
  timeout -k 10 300 python bench.py --rows 1250048 --steps 300 --warmup 10 --no-cpu-baseline --no-extras $variant \
    | python -c "import json,sys; d=json.loads([l for l in sys.stdin if l.startswith('{')][-1]); print(json.dumps({'variant': '$variant', 'ms_per_step': d['ms_per_step'], 'unchained': d.get('ms_per_step_unchained'), 'e2e_ms': d['e2e']['ms_per_step'], 'scan_ms': d['roofline']['kernel_ms'], 'dense_single_ms': d['extras']['dense_single_query']['ms']}))" \
    | tee -a "$OUT/shard_chained.jsonl"
done
timeout -k 10 300 python bench.py --steps 100 --warmup 5 --no-cpu-baseline --no-extras \
    | python -c "import json,sys; d=json.loads([l for l in sys.stdin if l.startswith('{')][-1]); print(json.dumps({'rows': 10000000, 'ms_per_step': d['ms_per_step'], 'unchained': d.get('ms_per_step_unchained'), 'e2e_ms': d['e2e']['ms_per_step'], 'scan_ms': d['roofline']['kernel_ms'], 'value': d['value']}))" \
    | tee -a "$OUT/shard_chained.jsonl"
timeout -k 10 300 python tools/cosine_bench.py 200000 | tee -a "$OUT/cosine.jsonl"
