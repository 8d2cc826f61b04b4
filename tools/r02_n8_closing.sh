#!/usr/bin/env bash
# Round-2 closing multi-GPU run on one 8-GPU box: correctness at 8 ranks, then the bench at N = 8, 4, 2.
set -u
OUT=gpurun_out/r02_n8_closing
mkdir -p "$OUT"
pick='import json,sys; d=json.loads([l for l in sys.stdin if l.startswith("{")][-1]); print(json.dumps({k: d.get(k) for k in ("n_gpus","value","ms_per_step","ms_per_step_unchained","gpu_launches","exchange")} | {"e2e": d["e2e"]["value"], "e2e_ms": d["e2e"]["ms_per_step"], "e2e_one_call": d["e2e"]["one_call_at_a_time"]["value"], "scan_ms": d["roofline"]["kernel_ms"], "frac": d["roofline"]["frac"], "traffic": d["roofline"]["traffic"], "rows_local": d["extras"]["dense_single_query"]["local_rows"]}))'
timeout -k 10 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29551 tools/multigpu_check.py 2>&1 | grep -v "^\*\|OMP_NUM" | tail -3 | tee "$OUT/multigpu_check.txt"
for n in 8 4 2; do
  timeout -k 10 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2956$n bench.py --gpus $n --steps 300 --warmup 10 2> "$OUT/bench_n$n.err" | tee "$OUT/bench_n$n.json" | python -c "$pick" | tee -a "$OUT/bench_summary.jsonl"
done
for f in "$OUT"/*.err; do tail -n 2 "$f" | cut -c1-300; done
