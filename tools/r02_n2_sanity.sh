#!/usr/bin/env bash
# two GPUs, closing sanity check of the tree: sharded == unsharded / chained steps, then the 2-GPU bench line as the driver launches it
set -u
OUT=gpurun_out/r02_n2_sanity
mkdir -p "$OUT"
timeout -k 10 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 tools/multigpu_check.py 2>&1 | tail -6 | tee "$OUT/multigpu_check.txt"
timeout -k 10 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29534 bench.py --gpus 2 --steps 100 --warmup 10 2> "$OUT/bench.err" | grep '^{' | tail -1 > "$OUT/bench_n2.json"
python -c "
import json; d=json.load(open('$OUT/bench_n2.json'))
print({k: d[k] for k in ('value','ms_per_step','n_gpus','gpu_launches','exchange') if k in d}, d['e2e']['value'])"
timeout -k 10 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29535 bench.py --impl reference --gpus 2 --steps 3 --warmup 1 2> "$OUT/ref.err" | grep '^{' | tail -1 | cut -c1-300
tail -3 "$OUT/bench.err"
