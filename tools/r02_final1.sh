#!/usr/bin/env bash
# Round-2 single-GPU closing run: the whole GPU suite, both bench arms, then the ncu launch list and the scan's traffic.
set -u
OUT=gpurun_out/r02_final
mkdir -p "$OUT"
timeout -k 10 1500 python -m pytest tests -m gpu -q 2>&1 | tail -15 | tee "$OUT/pytest_gpu.txt"
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2 | tee "$OUT/smoke.txt"
timeout -k 10 600 python bench.py --steps 20 --warmup 5 > "$OUT/bench_n1.json" 2> "$OUT/bench_n1.err"; echo "bench rc=$?"
timeout -k 10 600 python bench.py --impl reference --steps 20 --warmup 5 > "$OUT/bench_ref.json" 2> "$OUT/bench_ref.err"; echo "ref rc=$?"
# launch list of the bench command (extras and CPU baseline off: they are not the timed region)
python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-extras > "$OUT/plain.log" 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"flat_|bm25_|hybrid_|topk_|wide_|gemm_|cosine_|rows_to" -c 400 --csv --log-file "$OUT/launches_bench_n1.csv" \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-extras > "$OUT/ncu_launches.log" 2>&1
# dram traffic of the scan at the four shard sizes of the scaling run
for n in 10000000 5000064 2500096 1250048; do
  python tools/flat_profile.py 1 $n 2 > "$OUT/plain_flat_$n.log" 2>&1 && \
  ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,dram__throughput.avg.pct_of_peak_sustained_elapsed \
      --clock-control none -k regex:flat_scan_tma -s 1 -c 1 --csv --log-file "$OUT/traffic_$n.csv" \
      python tools/flat_profile.py 1 $n 2 > "$OUT/ncu_flat_$n.log" 2>&1
done
for args in "10000000 1" "10000000 64" "1250048 1"; do timeout 200 python tools/bm25_profile.py $args | tee -a "$OUT/bm25.jsonl"; done
B2R_BM25_BITS=64 timeout 200 python tools/bm25_profile.py 10000000 64 | tee -a "$OUT/bm25.jsonl"
ls -la "$OUT"
