#!/usr/bin/env bash
# Round-2 eight-GPU run: correctness at 8 ranks, the bench at N = 8 (and its r01-shaped variant) and N = 4,
# BASELINE config 2 with one year database per GPU, config 5 with the columns sharded over 8 GPUs.
set -u
OUT=gpurun_out/r02_n8
mkdir -p "$OUT"
G=$(nvidia-smi -L | wc -l)
echo "GPUs visible: $G"
pick='import json,sys; d=json.loads([l for l in sys.stdin if l.startswith("{")][-1]); print(json.dumps({k: d.get(k) for k in ("n_gpus","value","ms_per_step","ms_per_step_unchained","gpu_launches","exchange","steps_chained")} | {"e2e": d["e2e"]["value"], "e2e_ms": d["e2e"]["ms_per_step"], "scan_ms": d["roofline"]["kernel_ms"], "frac": d["roofline"]["frac"], "traffic": d["roofline"]["traffic"], "rows_local": d["extras"]["dense_single_query"]["local_rows"], "variant": sys.argv[1]}))'
timeout -k 10 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $G --master-addr 127.0.0.1 --master-port 29541 tools/multigpu_check.py 2>&1 | grep -v "^\*\|OMP_NUM" | tail -4 | tee "$OUT/multigpu_check.txt"
timeout -k 10 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $G --master-addr 127.0.0.1 --master-port 29542 bench.py --gpus $G --steps 400 --warmup 10 2> "$OUT/bench_n8.err" | tee "$OUT/bench_n8.json" | python -c "$pick" "" | tee -a "$OUT/bench_summary.jsonl"
timeout -k 10 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $G --master-addr 127.0.0.1 --master-port 29543 bench.py --gpus $G --steps 400 --warmup 10 --no-pipeline --tunable hybrid.no_fused_tail=1 2> "$OUT/bench_n8b.err" | python -c "$pick" "three kernels, un-chained (the r01 step shape)" | tee -a "$OUT/bench_summary.jsonl"
timeout -k 10 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29544 bench.py --gpus 4 --steps 300 --warmup 10 2> "$OUT/bench_n4.err" | tee "$OUT/bench_n4.json" | python -c "$pick" "" | tee -a "$OUT/bench_summary.jsonl"
timeout -k 10 400 python tools/yeardb_bench.py --reps 3 2> "$OUT/yeardb.err" | tail -1 | tee "$OUT/yeardb_bench.json" | cut -c1-1500
timeout -k 10 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $G --master-addr 127.0.0.1 --master-port 29545 tools/cosine_multi.py 200000 2> "$OUT/cosine_multi.err" | grep "^{" | tee "$OUT/cosine_multi.json"
tail -3 "$OUT"/*.err | cut -c1-300
