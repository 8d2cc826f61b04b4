#!/usr/bin/env bash
# fused tail (fusion / exchange inside the BM25 kernel's last CTA) + chained steps + K5 with a 4-stage ring
set -u
OUT=gpurun_out/r02_shard
mkdir -p "$OUT"
timeout -k 10 900 python -m pytest tests/test_hybrid_gpu.py tests/test_bm25_gpu.py tests/test_cosine_gpu.py tests/test_retrieval_gpu.py tests/test_config1.py -x -q 2>&1 | tail -6
for variant in "" "--no-pipeline" "--tunable hybrid.no_fused_tail=1"; do
  timeout -k 10 300 python bench.py --rows 1250048 --steps 400 --warmup 10 --no-cpu-baseline --no-extras $variant \
    | python -c "import json,sys; d=json.loads([l for l in sys.stdin if l.startswith('{')][-1]); print(json.dumps({'variant': '$variant', 'ms_per_step': round(d['ms_per_step'],4), 'unchained': round(d.get('ms_per_step_unchained'),4), 'e2e_ms': round(d['e2e']['ms_per_step'],4), 'scan_ms': round(d['roofline']['kernel_ms'],4), 'dense_single_ms': round(d['extras']['dense_single_query']['ms'],4), 'launches': d['gpu_launches']}))" \
    | tee -a "$OUT/shard_fused_tail.jsonl"
done
timeout -k 10 300 python bench.py --steps 100 --warmup 5 --no-cpu-baseline --no-extras \
    | python -c "import json,sys; d=json.loads([l for l in sys.stdin if l.startswith('{')][-1]); print(json.dumps({'rows': 10000000, 'ms_per_step': d['ms_per_step'], 'unchained': d.get('ms_per_step_unchained'), 'e2e_ms': d['e2e']['ms_per_step'], 'scan_ms': d['roofline']['kernel_ms'], 'value': d['value'], 'launches': d['gpu_launches']}))" \
    | tee -a "$OUT/shard_fused_tail.jsonl"
timeout -k 10 300 python tools/cosine_bench.py 200000 | tee -a "$OUT/cosine.jsonl"
