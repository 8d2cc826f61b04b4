#!/usr/bin/env bash
# two GPUs: correctness of the fused-tail exchange + chained steps, then the 2-GPU bench (and its un-chained / three-kernel variants)
set -u
OUT=gpurun_out/r02_n2
mkdir -p "$OUT"
if [ "${SKIP_CHECK:-0}" != "1" ]; then
timeout -k 10 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 tools/multigpu_check.py 2>&1 | tail -5 | tee "$OUT/multigpu_check.txt"
fi
for variant in "" "--no-pipeline" "--tunable hybrid.no_fused_tail=1 --no-pipeline"; do
  timeout -k 10 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29534 bench.py --gpus 2 --steps 300 --warmup 10 $variant 2> "$OUT/bench.err" \
    | python -c "import json,sys; d=json.loads([l for l in sys.stdin if l.startswith('{')][-1]); print(json.dumps({'variant': '$variant', 'value': round(d['value'],1), 'ms_per_step': round(d['ms_per_step'],4), 'unchained': round(d.get('ms_per_step_unchained'),4), 'e2e': round(d['e2e']['value'],1), 'e2e_ms': round(d['e2e']['ms_per_step'],4), 'scan_ms': round(d['roofline']['kernel_ms'],4), 'launches': d['gpu_launches'], 'exchange': d['exchange']}))" \
    | tee -a "$OUT/bench_n2.jsonl"
  tail -3 "$OUT/bench.err"
done
# the 8-GPU shard size on 2 GPUs (2 x 1.25M rows): the step the 8-GPU run will see, with a real exchange partner
for variant in "" "--no-pipeline" "--tunable hybrid.no_fused_tail=1 --no-pipeline"; do
  timeout -k 10 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29535 bench.py --gpus 2 --rows 2500096 --steps 500 --warmup 10 $variant 2> "$OUT/bench.err" \
    | python -c "import json,sys; d=json.loads([l for l in sys.stdin if l.startswith('{')][-1]); print(json.dumps({'rows': 2500096, 'variant': '$variant', 'value': round(d['value'],1), 'ms_per_step': round(d['ms_per_step'],4), 'unchained': round(d.get('ms_per_step_unchained'),4), 'e2e_ms': round(d['e2e']['ms_per_step'],4), 'scan_ms': round(d['roofline']['kernel_ms'],4), 'launches': d['gpu_launches']}))" \
    | tee -a "$OUT/bench_n2.jsonl"
  tail -3 "$OUT/bench.err"
done
