#!/usr/bin/env bash
# Round 2, single GPU: suite on the final kernels, k = 90 search timing, ncu captures of the step kernel, closing bench lines.
set -u
O=gpurun_out; mkdir -p $O
echo "== [1] pytest -m gpu"; timeout 1200 python -m pytest tests -q -m gpu 2>&1 | tail -6
echo "== [2] k=90 search, unsorted lists"
python - <<'PY'
import sys, torch
sys.path.insert(0, ".")
from bench import clustered
from torchdr_b200 import ops
ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
for n, d in ((1_000_000, 128), (1_000_000, 64)):
    X = clustered(n, d, torch.device("cuda"))
    for path, prune, q in (("tc", "on", n), ("tc", "off", 100_000)):
        for rep in range(2):
            torch.cuda.synchronize(); ev0.record()
            ops.knn(X[:q], X, 90, path=path, prune=prune)
            ev1.record(); torch.cuda.synchronize()
        print(f"k=90 {n}x{d} path={path} prune={prune} queries={q}: {ev0.elapsed_time(ev1):.1f} ms", flush=True)
    for rep in range(2):
        torch.cuda.synchronize(); ev0.record()
        ops.knn_umap_fused(X, X, 15, want_dist=False)
        ev1.record(); torch.cuda.synchronize()
    print(f"k=15 {n}x{d} fused pruned: {ev0.elapsed_time(ev1):.2f} ms", flush=True)
    del X
PY
echo "== [3] ncu --set full, persistent step kernel, 1M and 10M (5 iterations per launch)"
for pts in 1000000 10000000; do
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:umap_run_kernel_persist --launch-skip 2 -c 1 -f -o $O/r2_step_$pts \
      python bench.py --points $pts --steps 5 --warmup 3 --no-e2e --no-cpu --no-parity > $O/ncu_step_$pts.log 2>&1
  ncu -i $O/r2_step_$pts.ncu-rep --page raw --csv > $O/r2_step_${pts}_raw.csv 2>/dev/null
  python - $pts <<'PY'
import csv, sys
pts = sys.argv[1]
rows = list(csv.reader(open(f"gpurun_out/r2_step_{pts}_raw.csv", errors="replace")))
hdr, val = rows[0], rows[-1]
want = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "smsp__inst_executed.sum", "sm__inst_executed_pipe_xu.sum",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "dram__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active", "lts__t_sectors.sum", "lts__t_sector_hit_rate.pct",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "l1tex__t_sector_hit_rate.pct"]
for w in want:
    if w in hdr:
        print(pts, w, rows[1][hdr.index(w)] if len(rows) > 2 else "", val[hdr.index(w)])
PY
done
echo "== [4] closing bench lines (N = 1): default 10M, 1M, reference arm"
timeout 600 python bench.py --e2e-stages > $O/r2_bench_10m_n1.json 2> $O/r2_bench_10m_n1.err; tail -2 $O/r2_bench_10m_n1.err
timeout 300 python bench.py --points 1000000 --e2e-stages > $O/r2_bench_1m_n1.json 2> $O/r2_bench_1m_n1.err; tail -2 $O/r2_bench_1m_n1.err
timeout 300 python bench.py --impl reference --steps 20 --warmup 5 > $O/r2_bench_reference.json 2> $O/r2_bench_reference.err; tail -2 $O/r2_bench_reference.err
python - <<'PY'
import json
for f in ("r2_bench_10m_n1", "r2_bench_1m_n1", "r2_bench_reference"):
    try:
        d = json.loads(open(f"gpurun_out/{f}.json").read())
        e = d.get("e2e") or {}
        print(f, "value", round(d["value"], 3), "ms/step", round(d["ms_per_step"], 4), (d.get("roofline") or {}).get("frac"), "e2e", e.get("seconds") or e.get("value"),
              e.get("stages_seconds_rank0_instrumented_refit"), "knn", (d.get("affinity_kernel") or {}).get("ms"), (d.get("cpu_baseline") or {}).get("value_measured"), d.get("value_measured"))
    except Exception as ex:
        print(f, "FAILED", ex)
PY
