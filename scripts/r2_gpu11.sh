#!/usr/bin/env bash
set -u
O=gpurun_out; mkdir -p $O
echo "== [1] UMAP tests"; timeout 900 python -m pytest tests/test_gpu_parity.py -q -k "umap or estimators" 2>&1 | tail -3
echo "== [2] bench 1M / 10M (loop only)"
for pts in 1000000 10000000; do
  timeout 400 python bench.py --points $pts --no-e2e --no-cpu --no-parity --no-shuffled > $O/tail_$pts.json 2> $O/tail_$pts.err; tail -1 $O/tail_$pts.err
  python -c "
import json; d=json.loads(open('gpurun_out/tail_$pts.json').read()); print('$pts', 'value', round(d['value'],1), 'ms/step', round(d['ms_per_step'],4), 'frac', round(d['roofline']['frac'],3), d['clocks']['sm_mhz'])"
done
