#!/usr/bin/env bash
# Build libb2r.so (sm_100a only) in-tree.  Usage: ./build.sh [extra nvcc flags]
set -euo pipefail
cd "$(dirname "$0")"
PKG="rag-challenge-2_b200"
OUT="$PKG/libb2r.so"
SRCS=$(ls $PKG/csrc/*.cu)
nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 \
     -Xcompiler -fPIC -shared -Iinclude -I$PKG/csrc "$@" -o "$OUT" $SRCS
echo "built $OUT"
