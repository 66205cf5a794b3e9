#!/usr/bin/env bash
# Round-1 closing measurement on one B200 (run through gpurun): refreshes everything under profiles/ for the
# current default step kernel (umap_step_kernel_fast4).  Every leg has its own timeout and writes to gpurun_out/
# as it goes, so a clamped call still leaves the earlier legs.
set -u
O=gpurun_out
mkdir -p $O
NCU_STEP_METRICS=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__inst_executed.sum,smsp__thread_inst_executed_per_inst_executed.ratio,sm__warps_active.avg.pct_of_peak_sustained_active,smsp__issue_active.avg.pct_of_peak_sustained_active,lts__t_sectors.sum,lts__t_sector_hit_rate.pct,lts__t_sectors.avg.pct_of_peak_sustained_elapsed,gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed,sm__throughput.avg.pct_of_peak_sustained_elapsed,l1tex__t_sector_hit_rate.pct

echo "== [1] step kernel counters (3 launches)"; date +%T
timeout 400 ncu --metrics $NCU_STEP_METRICS --clock-control none -k regex:umap_step_kernel_fast4 -s 30 -c 3 --csv \
  --log-file $O/step_counters.csv python bench.py --steps 40 --warmup 5 --no-e2e --no-cpu > $O/step_counters.log 2>&1
python scripts/ncu_summary.py traffic $O/step_counters.csv $O/step_kernel_traffic.json umap_step_kernel_fast4 \
  && cp $O/step_kernel_traffic.json profiles/step_kernel_traffic.json

echo "== [2] bench (default flags)"; date +%T
timeout 600 python bench.py > $O/r1_bench.json 2> $O/r1_bench.err
tail -c 600 $O/r1_bench.json; echo

echo "== [3] launch list"; date +%T
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file $O/r1_launches.csv \
  python bench.py --steps 50 --warmup 5 --no-cpu > $O/launches.log 2>&1
python scripts/ncu_summary.py launches $O/r1_launches.csv $O/r1_launches_summary.txt

echo "== [4] ncu --set full, one launch of the step kernel"; date +%T
timeout 400 ncu --set full --import-source on --clock-control none -k regex:umap_step_kernel_fast4 -s 30 -c 1 -f \
  -o $O/step_fast4_full python bench.py --steps 40 --warmup 5 --no-e2e --no-cpu > $O/step_full.log 2>&1
ls -la $O/*.ncu-rep

echo "== [5] pytest -m gpu"; date +%T
timeout 900 python -m pytest tests -m gpu -x -q > $O/pytest_gpu.log 2>&1
tail -5 $O/pytest_gpu.log

echo "== [6] reference arm"; date +%T
timeout 300 python bench.py --impl reference --steps 20 --warmup 3 > $O/r1_bench_reference.json 2> $O/r1_bench_reference.err
tail -c 400 $O/r1_bench_reference.json; echo

echo "== [7] 10 M x 128 (north-star size), loop only"; date +%T
timeout 500 python bench.py --points 10000000 --steps 300 --warmup 10 --no-e2e --no-cpu > $O/r1_bench_10m.json 2> $O/r1_bench_10m.err
tail -c 400 $O/r1_bench_10m.json; echo
date +%T
