#!/usr/bin/env bash
set -uo pipefail
mkdir -p gpurun_out/r02_bm25
timeout 900 python -m pytest tests/test_bm25_gpu.py tests/test_published_vectors_gpu.py tests/test_random_shapes_gpu.py tests/test_hybrid_gpu.py tests/test_retrieval_gpu.py -q -k "bm25 or hybrid or published" 2>&1 | tail -15
{
timeout 300 python tools/bm25_profile.py 10000000 64
timeout 300 python tools/bm25_profile.py 10000000 1
timeout 300 python tools/bm25_profile.py 10000000 4
timeout 300 python tools/bm25_profile.py 10000000 4 bm25.pilot_blocks=-1
timeout 300 python tools/bm25_profile.py 10000000 16
timeout 300 python tools/bm25_profile.py 10000000 16 bm25.pilot_blocks=-1
timeout 300 python tools/bm25_profile.py 1250000 1
timeout 300 python tools/bm25_profile.py 1250000 64
} 2>&1 | grep '^{' | tee gpurun_out/r02_bm25/slots_$(date +%H%M%S).jsonl
