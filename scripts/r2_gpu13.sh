#!/usr/bin/env bash
# A/B of the candidate-list structure (max-heap vs unsorted list + rescan) on one box.
# Needs two prebuilt libraries (build/ travels to the box): build/libtdrb200_heap.so = ./build.sh of the tree under test,
# build/libtdrb200_scan.so = the same link line with knn_tc.o compiled from the revision to compare against.
set -uo pipefail
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -3 | tee gpurun_out/r2_pytest_gpu_heap.log
for v in heap scan; do
  cp build/libtdrb200_$v.so torchdr_b200/lib/libtdrb200.so
  echo "== variant $v"
  python scripts/knn_time.py 1000000 128 15 generator 2>&1 | grep prune
  python scripts/knn_time.py 1000000 128 90 generator 2>&1 | grep prune
  python scripts/knn_time.py 10000000 128 15 generator 2>&1 | grep prune
  python scripts/knn_time.py 4000000 64 90 generator 2>&1 | grep prune
done 2>&1 | tee gpurun_out/r2_heap_ab.txt
