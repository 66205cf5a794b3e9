set -x
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -8
for cfg in "1 5" "2 4" "2 5" "2 6" "2 8"; do
  set -- $cfg
  echo "== variant $1 occ $2"
  TDR_STEP_FAST=$1 TDR_STEP_OCC=$2 timeout 300 python bench.py --steps 500 --warmup 20 --no-e2e --no-cpu 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('it/s', round(d['value'],1), 'ms', round(d['ms_per_step'],4), 'frac', d['roofline']['frac'])"
done
for v in 1 2; do
TDR_STEP_FAST=$v timeout 600 ncu --metrics smsp__inst_executed.sum,gpu__time_duration.sum,smsp__thread_inst_executed_per_inst_executed.ratio,sm__warps_active.avg.pct_of_peak_sustained_active,smsp__issue_active.avg.pct_of_peak_sustained_active --clock-control none -k regex:umap_step_kernel_fast -s 30 -c 2 --csv --log-file gpurun_out/step_inst_v$v.csv python bench.py --steps 40 --warmup 5 --no-e2e --no-cpu > /dev/null 2>&1
tail -12 gpurun_out/step_inst_v$v.csv
done
