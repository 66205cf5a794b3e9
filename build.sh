#!/usr/bin/env bash
# Build libtdrb200.so (sm_100a only) in-tree.  Usage: ./build.sh [extra nvcc flags]
set -euo pipefail
ROOT="$(cd "$(dirname "$0")" && pwd)"
SRC="$ROOT/torchdr_b200/csrc"
OUT="$ROOT/torchdr_b200/lib"
NVCC="${NVCC:-/usr/local/cuda/bin/nvcc}"
mkdir -p "$OUT" "$ROOT/build"
FLAGS=(-O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -Xcompiler -fPIC
       -Xcompiler -fvisibility=hidden --expt-relaxed-constexpr "$@")
pids=()
for f in capi knn knn_tc affinity graph umap_step autograd_steps indexed reorder; do
  "$NVCC" "${FLAGS[@]}" -c "$SRC/$f.cu" -o "$ROOT/build/$f.o" &
  pids+=($!)
done
for p in "${pids[@]}"; do wait "$p"; done
"$NVCC" -shared -gencode arch=compute_100a,code=sm_100a -o "$OUT/libtdrb200.so" \
  "$ROOT"/build/{capi,knn,knn_tc,affinity,graph,umap_step,autograd_steps,indexed,reorder}.o -lcudart_static -ldl -lrt -lpthread
echo "built $OUT/libtdrb200.so"
