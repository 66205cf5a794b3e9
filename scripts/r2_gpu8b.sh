#!/usr/bin/env bash
# Round 2, closing 8-GPU run on the final code: sharded-path check + the default bench line.
set -u
N=${1:-8}
O=gpurun_out; mkdir -p $O
echo "== [1] dist_check x$N"
timeout 420 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29611 scripts/dist_check.py 2>&1 | grep -E "dist_check|Error|error" | tail -40
echo "== [2] bench 10M x$N"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29612 bench.py --gpus $N --steps 20 --warmup 5 --e2e-stages > $O/r2_bench_10m_n$N.json 2> $O/r2_bench_10m_n$N.err; tail -3 $O/r2_bench_10m_n$N.err
python - $N <<'PY'
import json, sys
N = sys.argv[1]
d = json.loads(open(f"gpurun_out/r2_bench_10m_n{N}.json").read())
e = d["e2e"]
print("value", round(d["value"], 1), "ms/step", round(d["ms_per_step"], 4), "frac", round(d["roofline"]["frac"], 3), "e2e s", round(e["seconds"], 3), e["exchange"],
      e["stages_seconds_rank0_instrumented_refit"], "shuffled", e["shuffled_rows"], "knn ms", d["affinity_kernel"]["ms"], "blocks", [round(x, 2) for x in d["timing"]["block_ms_max_over_ranks"]], d["clocks"], d["parity"]["sigma_rows_vs_oracle"])
PY
