#!/usr/bin/env bash
# Config 2 host path after the prefetching record helper + K1 at 1..8 queries over 10M rows.
set -uo pipefail
mkdir -p gpurun_out/r02_host
python -m pytest tests/test_retrieval_gpu.py tests/test_host_records.py -x -q 2>&1 | tail -3
rm -f gpurun_out/yeardb_bench.json
for v in "--python-records" "" "--lazy-records"; do
  timeout 600 python tools/yeardb_bench.py --reps 7 $v 2>&1 | tail -1 > /tmp/line.json
  python - "$v" <<'PY'
import json, sys
d = json.loads(open("/tmp/line.json").read())
m = d["mixes"]
print(json.dumps({"variant": sys.argv[1], "chunk_records": d["chunk_records"], "record_helper": d["record_helper"],
                  **{k: {"batch_ms": round(v["batch_s"] * 1e3, 2), "device_ms": v["device_search_ms_max_over_gpus"],
                         "loop_ms": round(v["loop_s"] * 1e3, 1), "same": v["batch_equals_loop"]} for k, v in m.items()}}))
PY
done | tee gpurun_out/r02_host/yeardb_variants2.jsonl
cp gpurun_out/yeardb_bench.json gpurun_out/r02_host/yeardb_full2.jsonl
timeout 300 python tools/flat_sweep.py 10000000 1024 defaults | tee gpurun_out/r02_host/k1_small_batches.jsonl
timeout 300 python tools/flat_sweep.py 10000000 1024 defaults flat.gemm_min_nq=2 | tee gpurun_out/r02_host/k1_small_batches_gemm2.jsonl
