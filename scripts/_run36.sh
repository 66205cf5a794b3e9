echo "== umap tests with TDR_STEP_FAST=4 cfg 7 (PIPE + CG)"
TDR_STEP_FAST=4 TDR_STEP_CFG=7 timeout 600 python -m pytest tests -m gpu -x -q -k "umap or estimators or long_run" 2>&1 | tail -4
for cfg in 0 1 2 3 4 6 7 8 9; do
  echo "== fast4 cfg $cfg"
  TDR_STEP_FAST=4 TDR_STEP_CFG=$cfg timeout 300 python bench.py --steps 400 --warmup 20 --no-e2e --no-cpu 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('it/s', round(d['value'],1), 'ms', round(d['ms_per_step'],4), 'frac', round(d['roofline']['frac'],4))"
done
