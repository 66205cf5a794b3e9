#!/usr/bin/env bash
set -u
O=gpurun_out; mkdir -p $O
echo "== [1] tree-order tests"; timeout 600 python -m pytest tests/test_gpu_parity.py -q -s -k "tree_order or certified or estimators" 2>&1 | tail -6
echo "== [2] tree order timing"
timeout 300 python scripts/knn_time.py 1000000 128 15 tree 2>&1 | tail -5
timeout 300 python scripts/knn_time.py 10000000 128 15 tree 2>&1 | tail -4
timeout 300 python scripts/e2e_breakdown.py 1000000 shuffled 2>&1 | tail -2
