echo "== full suite (default variant 3)"
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
echo "== umap tests with TDR_STEP_FAST=4"
TDR_STEP_FAST=4 timeout 600 python -m pytest tests -m gpu -x -q -k "umap or estimators or long_run" 2>&1 | tail -6
for cfg in "4 4" "4 3" "3 4"; do
  set -- $cfg
  echo "== variant $1 occ $2"
  TDR_STEP_FAST=$1 TDR_STEP_OCC=$2 timeout 300 python bench.py --steps 500 --warmup 20 --no-e2e --no-cpu 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('it/s', round(d['value'],1), 'ms', round(d['ms_per_step'],4), 'frac', d['roofline']['frac'])"
done
TDR_STEP_FAST=4 timeout 600 ncu --metrics smsp__inst_executed.sum,gpu__time_duration.sum,smsp__thread_inst_executed_per_inst_executed.ratio,sm__warps_active.avg.pct_of_peak_sustained_active,smsp__issue_active.avg.pct_of_peak_sustained_active,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sectors.sum,l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum --clock-control none -k regex:umap_step_kernel_fast -s 30 -c 1 --csv --log-file gpurun_out/step_inst_v4.csv python bench.py --steps 40 --warmup 5 --no-e2e --no-cpu > /dev/null 2>&1
tail -10 gpurun_out/step_inst_v4.csv | cut -d, -f5,13-
