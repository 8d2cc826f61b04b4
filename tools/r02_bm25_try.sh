#!/usr/bin/env bash
# BM25 streaming kernel: parity tests, then the timing points of profiles/r02_bm25.md.
set -uo pipefail
mkdir -p gpurun_out/r02_bm25
python -m pytest tests/test_bm25_gpu.py tests/test_published_vectors_gpu.py -x -q 2>&1 | tail -3
{
timeout 300 python tools/bm25_profile.py 10000000 64
timeout 300 python tools/bm25_profile.py 10000000 64 bm25.cfg=1
timeout 300 python tools/bm25_profile.py 10000000 8
timeout 300 python tools/bm25_profile.py 10000000 1
B2R_BM25_BITS=64 timeout 300 python tools/bm25_profile.py 10000000 64
} 2>&1 | grep '^{' | tee gpurun_out/r02_bm25/try_$(date +%H%M%S).jsonl
