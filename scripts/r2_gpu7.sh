#!/usr/bin/env bash
# Round 2, single GPU: suite + smoke on the final kernels, kNN timings after the grouped candidate scan, default bench line.
set -u
O=gpurun_out; mkdir -p $O
echo "== [1] smoke + pytest -m gpu"; timeout 300 python __graft_entry__.py smoke 2>&1 | tail -2; timeout 1200 python -m pytest tests -q -m gpu 2>&1 | tail -4
echo "== [2] kNN timings"
for o in generator tree; do timeout 300 python scripts/knn_time.py 1000000 128 15 $o 2>&1 | tail -4; done
timeout 600 python scripts/knn_time.py 10000000 128 15 generator 2>&1 | tail -3
timeout 300 python scripts/knn_time.py 1000000 128 90 generator 2>&1 | tail -4
echo "== [3] default bench (10M, with shuffled legs)"
timeout 900 python bench.py > $O/r2_bench_10m_n1.json 2> $O/r2_bench_10m_n1.err; tail -2 $O/r2_bench_10m_n1.err
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r2_bench_10m_n1.json").read())
a, e = d["affinity_kernel"], d["e2e"]
print("value", round(d["value"], 1), "frac", round(d["roofline"]["frac"], 3), "traffic", d["roofline"]["traffic"], "knn ms", round(a["ms"], 1), "shuffled", a["shuffled_rows"],
      "full sweep tflops", round(a["full_sweep_tflops_2nnd"], 1), "e2e", round(e["seconds"], 3), "e2e shuffled", e["shuffled_rows"], "cpu", d["cpu_baseline"]["value"], d["cpu_baseline"]["value_measured"], d["clocks"])
PY
