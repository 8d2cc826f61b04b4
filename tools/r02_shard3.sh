#!/usr/bin/env bash
# Where do the ~30 us between the dense scan alone and the hybrid step go (8-GPU shard on one GPU)?
set -u
OUT=gpurun_out/r02_shard
mkdir -p "$OUT"
timeout -k 10 900 python -m pytest tests/test_hybrid_gpu.py tests/test_bm25_gpu.py tests/test_flat_ip_gpu.py tests/test_retrieval_gpu.py -x -q 2>&1 | tail -4
for variant in "" "--fake-bm25" "--fake-bm25 --no-pipeline" "--tunable bm25.debug_skip=3" "--tunable bm25.v1=2" "--no-overlap"; do
  timeout -k 10 300 python bench.py --rows 1250048 --steps 400 --warmup 10 --no-cpu-baseline --no-extras $variant \
    | python -c "import json,sys; d=json.loads([l for l in sys.stdin if l.startswith('{')][-1]); print(json.dumps({'variant': '$variant', 'ms_per_step': round(d['ms_per_step'],4), 'unchained': round(d.get('ms_per_step_unchained'),4), 'e2e_ms': round(d['e2e']['ms_per_step'],4), 'scan_ms': round(d['roofline']['kernel_ms'],4), 'dense_single_ms': round(d['extras']['dense_single_query']['ms'],4)}))" \
    | tee -a "$OUT/shard_where.jsonl"
done
