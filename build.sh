#!/usr/bin/env bash
# Build libb2r.so (sm_100a only) in-tree.  Usage: ./build.sh [extra nvcc flags]
# Every csrc/*.cu is compiled to an object in parallel, then linked into one shared library.
set -euo pipefail
cd "$(dirname "$0")"
PKG="rag-challenge-2_b200"
OUT="$PKG/libb2r.so"
OBJ="build/obj"
mkdir -p "$OBJ"
FLAGS=(-gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC -Iinclude -I$PKG/csrc "$@")
pids=()
objs=()
for src in $PKG/csrc/*.cu; do
  o="$OBJ/$(basename "${src%.cu}").o"
  objs+=("$o")
  nvcc "${FLAGS[@]}" -c "$src" -o "$o" &
  pids+=($!)
done
for p in "${pids[@]}"; do wait "$p"; done
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o "$OUT" "${objs[@]}"
echo "built $OUT"
