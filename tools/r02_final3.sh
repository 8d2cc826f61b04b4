#!/usr/bin/env bash
# Round-2 closing run, final tree (BM25 ring geometry / select scratch inside the accumulators): the whole GPU suite, smoke, both bench arms, the launch list
# of the bench command, one full ncu capture of the BM25 streaming kernel's main launch.
set -u
OUT=gpurun_out/r02_final3
mkdir -p "$OUT"
timeout -k 10 1500 python -m pytest tests -m gpu -q 2>&1 | tail -15 | tee "$OUT/pytest_gpu.txt"
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2 | tee "$OUT/smoke.txt"
timeout -k 10 900 python bench.py --steps 20 --warmup 5 > "$OUT/bench_n1.json" 2> "$OUT/bench_n1.err"; echo "bench rc=$?"
timeout -k 10 600 python bench.py --impl reference --steps 20 --warmup 5 > "$OUT/bench_ref.json" 2> "$OUT/bench_ref.err"; echo "ref rc=$?"
for args in "10000000 1" "10000000 8" "10000000 64" "1250048 1"; do timeout 200 python tools/bm25_profile.py $args | grep '^{' | tee -a "$OUT/bm25.jsonl"; done
B2R_BM25_BITS=64 timeout 200 python tools/bm25_profile.py 10000000 64 | grep '^{' | tee -a "$OUT/bm25.jsonl"
# launches of the streaming kernel alternate pilot / main: the fourth one is a main launch
timeout 600 ncu --set full --clock-control none --import-source on -k regex:bm25_stream -s 3 -c 1 -f -o "$OUT/prof_bm25s_final" \
    python tools/bm25_profile.py 10000000 64 > "$OUT/ncu_bm25.log" 2>&1
ncu -i "$OUT/prof_bm25s_final.ncu-rep" --page details --csv > "$OUT/prof_bm25s_final_details.csv" 2>/dev/null
ncu -i "$OUT/prof_bm25s_final.ncu-rep" --page raw --csv > "$OUT/prof_bm25s_final_raw.csv" 2>/dev/null
ls -la "$OUT"

