TDR_STEP_CFG=5 timeout 200 python -m pytest tests/test_gpu_parity.py -x -q -k "umap" 2>&1 | tail -2
for cfg in 0 5; do
  TDR_STEP_CFG=$cfg timeout 200 python bench.py --points 10000000 --steps 200 --warmup 10 --no-e2e --no-cpu 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('10M cfg $cfg it/s', round(d['value'],1), 'ms', round(d['ms_per_step'],4), d['clocks']['sm_mhz'], d['clocks']['reasons'])"
done
TDR_STEP_CFG=5 timeout 200 python bench.py --steps 1000 --warmup 10 --no-e2e --no-cpu 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('1M cfg 5 it/s', round(d['value'],1), 'ms', round(d['ms_per_step'],4))"
