#!/usr/bin/env bash
set -u
echo "== kNN timings (64-bit key lists, no back-off)"
timeout 300 python scripts/knn_time.py 1000000 128 15 generator 2>&1 | tail -4
timeout 300 python scripts/knn_time.py 1000000 128 90 generator 2>&1 | tail -4
timeout 300 python scripts/knn_time.py 1000000 128 90 generator 2>&1 | tail -4
timeout 600 python scripts/knn_time.py 10000000 128 15 generator 2>&1 | tail -3
