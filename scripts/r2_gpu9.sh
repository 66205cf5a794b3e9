#!/usr/bin/env bash
# Round 2, one B200: ncu capture of the pruned sweep's phase B at 10 M, LargeVis config on the final kNN kernel.
set -u
O=gpurun_out; mkdir -p $O
echo "== [1] ncu --set full, knn_tc_kernel phase B at 10M"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:knn_tc_kernel --launch-skip 1 -c 1 -f -o $O/r2_knn_phaseB_10m \
    python scripts/knn_time.py 10000000 128 15 generator > $O/ncu_knn.log 2>&1
ncu -i $O/r2_knn_phaseB_10m.ncu-rep --page raw --csv > $O/r2_knn_phaseB_10m_raw.csv 2>/dev/null
python - <<'PY'
import csv
rows = list(csv.reader(open("gpurun_out/r2_knn_phaseB_10m_raw.csv", errors="replace")))
hdr, val = rows[0], rows[-1]
for w in ["gpu__time_duration.sum", "dram__bytes_read.sum", "smsp__inst_executed.sum", "sm__inst_executed_pipe_tensor.sum", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
          "smsp__issue_active.avg.pct_of_peak_sustained_active", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
          "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__grid_size",
          "smsp__average_warp_latency_issue_stalled_long_scoreboard.ratio", "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
          "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio", "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
          "smsp__average_warps_issue_stalled_membar_per_issue_active.ratio", "smsp__average_warps_issue_stalled_sleeping_per_issue_active.ratio"]:
    if w in hdr:
        print(w, val[hdr.index(w)])
PY
ncu -i $O/r2_knn_phaseB_10m.ncu-rep --page source --csv 2>/dev/null > $O/r2_knn_phaseB_10m_src.csv
python - <<'PY'
import csv
rows = list(csv.reader(open("gpurun_out/r2_knn_phaseB_10m_src.csv", errors="replace")))
hdr = rows[1]; data = rows[2:]
ix = {h: i for i, h in enumerate(hdr)}
tot = sum(int(r[ix["# Samples"]] or 0) for r in data)
print("stall samples", tot, "warp inst", sum(int(r[ix["Instructions Executed"]] or 0) for r in data))
for r in sorted(data, key=lambda r: -int(r[ix["# Samples"]] or 0))[:14]:
    print(f'{100.0 * int(r[ix["# Samples"]]) / tot:5.1f} %  {int(r[ix["Instructions Executed"]]):>11}  {r[ix["Source"]].strip()[:90]}')
PY
