#!/usr/bin/env bash
set -u
echo "== [1] pytest -m gpu"; timeout 1200 python -m pytest tests -q -m gpu 2>&1 | tail -4
echo "== [2] kNN timings (tile-or-atom ring stages, key lists, two sigma rows per warp)"
timeout 300 python scripts/knn_time.py 1000000 128 15 generator 2>&1 | tail -4
timeout 300 python scripts/knn_time.py 1000000 128 90 generator 2>&1 | tail -4
timeout 600 python scripts/knn_time.py 10000000 128 15 generator 2>&1 | tail -3
timeout 300 python scripts/knn_time.py 1000000 256 15 generator 2>&1 | tail -4
