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
# host-side CPython helper (record dictionaries of the batched retrieval path); plain C, no CUDA.  Optional: without it
# retrieval.py runs the same loops in Python, so a box without Python.h still gets a working library.
PYINC="$(python -c 'import sysconfig; print(sysconfig.get_paths()["include"])' 2>/dev/null || true)"
if [ -n "$PYINC" ] && [ -f "$PYINC/Python.h" ] && \
   gcc -O2 -fPIC -shared -Wall -I"$PYINC" "$PKG/csrc/host_records.c" -o "$PKG/_b2r_records.so" -lm; then
  echo "built $PKG/_b2r_records.so"
else
  echo "warning: $PKG/_b2r_records.so not built (no Python.h or gcc failed): record dictionaries will be built in Python" >&2
fi
