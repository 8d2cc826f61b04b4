#!/usr/bin/env bash
set -uo pipefail
mkdir -p gpurun_out/r02_bm25
{
for p in 0 32 64 128 256; do timeout 200 python tools/bm25_profile.py 10000000 1 bm25.pilot_blocks=$p; done
for p in 0 32 64; do timeout 200 python tools/bm25_profile.py 10000000 2 bm25.pilot_blocks=$p; done
for p in 0 16 32; do timeout 200 python tools/bm25_profile.py 1250048 1 bm25.pilot_blocks=$p; done
} 2>&1 | grep '^{' | tee gpurun_out/r02_bm25/pilot1_$(date +%H%M%S).jsonl
