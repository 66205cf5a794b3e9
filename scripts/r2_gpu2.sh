#!/usr/bin/env bash
# Round 2, N GPUs (default 2): remaining GPU tests, sharded-path check, bench at 10 M and 1 M.
set -u
N=${1:-2}
O=gpurun_out; mkdir -p $O
echo "== [1] pytest -m gpu (rest)"; timeout 900 python -m pytest tests -q -m gpu 2>&1 | tail -8
echo "== [2] dist_check x$N"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29611 scripts/dist_check.py 2>&1 | grep -E "dist_check|Error|error" | tail -40
echo "== [3] bench 10M x$N"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29612 bench.py --gpus $N --steps 20 --warmup 5 > $O/r2_bench_10m_n$N.json 2> $O/r2_bench_10m_n$N.err; tail -3 $O/r2_bench_10m_n$N.err
echo "== [4] bench 1M x$N"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29613 bench.py --gpus $N --steps 20 --warmup 5 --points 1000000 > $O/r2_bench_1m_n$N.json 2> $O/r2_bench_1m_n$N.err; tail -3 $O/r2_bench_1m_n$N.err
python - $N <<'PY'
import json, sys
N = sys.argv[1]
for f in (f"gpurun_out/r2_bench_10m_n{N}.json", f"gpurun_out/r2_bench_1m_n{N}.json"):
    try:
        d = json.loads(open(f).read())
        print(f, "value", round(d["value"], 1), "ms/step", round(d["ms_per_step"], 4), "frac", round(d["roofline"]["frac"], 3),
              "e2e s", round(d["e2e"]["seconds"], 3), d["e2e"].get("exchange"), "knn ms", round(d["affinity_kernel"]["ms"], 2),
              "blocks", [round(x, 2) for x in d["timing"]["block_ms_max_over_ranks"]], "clocks", d["clocks"]["sm_mhz"], d["clocks"]["reasons"])
    except Exception as e:
        print(f, "FAILED", e)
PY
