#!/usr/bin/env bash
# Final list structure (heap for k > 32, unsorted + re-scan below): suite, kNN timings, configs[3] on one GPU.
set -uo pipefail
O=gpurun_out; mkdir -p $O
python -m pytest tests -m gpu -x -q 2>&1 | tail -3 | tee $O/r2_pytest_gpu.log
{
  python scripts/knn_time.py 1000000 128 15 generator 2>&1 | grep prune
  python scripts/knn_time.py 1000000 128 90 generator 2>&1 | grep prune
  python scripts/knn_time.py 4000000 64 90 generator 2>&1 | grep prune
} 2>&1 | tee $O/r2_heap_final.txt
timeout 300 python bench.py --config c4 --steps 10 --warmup 5 > $O/r2_c4_n1.json 2> $O/r2_c4_n1.err; tail -2 $O/r2_c4_n1.err
python - <<'PY'
import json
d = json.load(open("gpurun_out/r2_c4_n1.json"))
print("c4 n1: value", round(d["value"], 1), "ms", round(d["ms_per_step"], 3), "e2e", d["e2e"].get("seconds"), d.get("affinity_kernel"))
PY
