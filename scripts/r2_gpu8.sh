#!/usr/bin/env bash
# Round 2, 8 GPUs: sharded-path check, bench at the north-star size, BASELINE configs[3] (LargeVis 10M x 64) and [4] (UMAP 50M x 96).
set -u
N=${1:-8}
O=gpurun_out; mkdir -p $O
run() { python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $1 "${@:2}"; }
echo "== [1] dist_check x$N"
timeout 420 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29611 scripts/dist_check.py 2>&1 | grep -E "dist_check|Error|error" | tail -40
echo "== [2] bench 10M x$N"
timeout 420 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29612 bench.py --gpus $N --steps 20 --warmup 5 --e2e-stages > $O/r2_bench_10m_n$N.json 2> $O/r2_bench_10m_n$N.err; tail -3 $O/r2_bench_10m_n$N.err
echo "== [3] c4 LargeVis 10M x 64 x$N"
timeout 420 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29613 bench.py --gpus $N --config c4 --steps 10 --warmup 5 > $O/r2_c4_n$N.json 2> $O/r2_c4_n$N.err; tail -3 $O/r2_c4_n$N.err
echo "== [4] c5 UMAP 50M x 96 x$N"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29614 bench.py --gpus $N --config c5 --steps 10 --warmup 5 --no-e2e > $O/r2_c5_n$N.json 2> $O/r2_c5_n$N.err; tail -3 $O/r2_c5_n$N.err
python - $N <<'PY'
import json, sys
N = sys.argv[1]
for f in (f"gpurun_out/r2_bench_10m_n{N}.json", f"gpurun_out/r2_c4_n{N}.json", f"gpurun_out/r2_c5_n{N}.json"):
    try:
        d = json.loads(open(f).read())
        e = d.get("e2e") or {}
        print(f, "value", round(d["value"], 1), "ms/step", round(d["ms_per_step"], 4), "frac", round(d["roofline"]["frac"], 3),
              "e2e s", e.get("seconds"), e.get("exchange"), e.get("stages_seconds_rank0_instrumented_refit"),
              "knn ms", (d.get("affinity_kernel") or {}).get("ms"), "aff s", d.get("affinity_seconds"), d.get("reference_formulation_ms_per_iteration"),
              "blocks", [round(x, 2) for x in d["timing"]["block_ms_max_over_ranks"]], "clocks", d["clocks"]["sm_mhz"], d["clocks"]["reasons"],
              (d.get("parity") or {}).get("knn_sampled_rows_fp64"))
    except Exception as ex:
        print(f, "FAILED", ex)
PY
