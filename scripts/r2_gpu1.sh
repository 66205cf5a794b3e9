#!/usr/bin/env bash
# Round 2, single GPU: GPU suite, bench at 1 M and 10 M (persistent step kernel), kNN on shuffled rows.
set -u
O=gpurun_out; mkdir -p $O
echo "== [1] pytest -m gpu"; timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -15
echo "== [2] bench 1M"; timeout 300 python bench.py --points 1000000 --steps 20 --warmup 5 --no-cpu > $O/r2_bench_1m.json 2> $O/r2_bench_1m.err; tail -3 $O/r2_bench_1m.err
python - <<'PY'
import json
for f in ("gpurun_out/r2_bench_1m.json",):
    try:
        d = json.loads(open(f).read())
        print(f, "value", round(d["value"], 1), "ms/step", round(d["ms_per_step"], 4), "frac", round(d["roofline"]["frac"], 3),
              "e2e s", round(d["e2e"]["seconds"], 3), "knn ms", round(d["affinity_kernel"]["ms"], 2), "blocks", [round(x, 2) for x in d["timing"]["block_ms_max_over_ranks"]],
              "parity", d["parity"]["knn_sampled_rows_fp64"])
    except Exception as e:
        print(f, "FAILED", e)
PY
echo "== [3] bench 10M (default)"; timeout 600 python bench.py > $O/r2_bench_10m.json 2> $O/r2_bench_10m.err; tail -3 $O/r2_bench_10m.err
python - <<'PY'
import json
for f in ("gpurun_out/r2_bench_10m.json",):
    try:
        d = json.loads(open(f).read())
        print(f, "value", round(d["value"], 1), "ms/step", round(d["ms_per_step"], 4), "frac", round(d["roofline"]["frac"], 3),
              "e2e s", round(d["e2e"]["seconds"], 3), "knn ms", round(d["affinity_kernel"]["ms"], 2), "blocks", [round(x, 2) for x in d["timing"]["block_ms_max_over_ranks"]],
              "clocks", d["clocks"], "cpu", d.get("cpu_baseline", {}).get("value"), "parity", d["parity"]["knn_sampled_rows_fp64"])
    except Exception as e:
        print(f, "FAILED", e)
PY
echo "== [4] kNN by row order, 1M"
for o in generator shuffled tree; do timeout 300 python scripts/knn_time.py 1000000 128 15 $o 2>&1 | tail -6; done
echo "== [5] e2e shuffled 1M (auto re-order)"
timeout 300 python bench.py --points 1000000 --order shuffled --steps 20 --no-cpu > $O/r2_bench_1m_shuffled.json 2> $O/r2_bench_1m_shuffled.err; tail -2 $O/r2_bench_1m_shuffled.err
python -c "
import json; d=json.loads(open('gpurun_out/r2_bench_1m_shuffled.json').read()); print('shuffled 1M: value', round(d['value'],1), 'e2e s', round(d['e2e']['seconds'],3), 'bench kNN ms', round(d['affinity_kernel']['ms'],1))"
