#!/usr/bin/env bash
set -uo pipefail
mkdir -p gpurun_out/r02_bm25
timeout 400 python -m pytest tests/test_bm25_gpu.py tests/test_published_vectors_gpu.py tests/test_random_shapes_gpu.py -x -q -k "bm25 or hybrid or published" 2>&1 | tail -4
{
timeout 100 python tools/bm25_profile.py 10000000 64
B2R_BM25_BITS=64 timeout 100 python tools/bm25_profile.py 10000000 64
timeout 100 python tools/bm25_profile.py 10000000 1
} 2>&1 | grep '^{' | tee gpurun_out/r02_bm25/probe_$(date +%H%M%S).jsonl
