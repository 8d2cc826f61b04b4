#!/usr/bin/env bash
set -uo pipefail
mkdir -p gpurun_out/r02_bm25
{
timeout 200 python tools/bm25_profile.py 10000000 64
timeout 200 python tools/bm25_profile.py 10000000 64 bm25.cfg=3
timeout 200 python tools/bm25_profile.py 10000000 64 bm25.cfg=4
timeout 200 python tools/bm25_profile.py 10000000 64 bm25.pilot_blocks=14
B2R_BM25_BITS=64 timeout 200 python tools/bm25_profile.py 10000000 64
B2R_BM25_BITS=64 timeout 200 python tools/bm25_profile.py 10000000 64 bm25.cfg=3
B2R_BM25_BITS=64 timeout 200 python tools/bm25_profile.py 10000000 64 bm25.cfg=4
} 2>&1 | grep '^{' | tee gpurun_out/r02_bm25/geom2_$(date +%H%M%S).jsonl
