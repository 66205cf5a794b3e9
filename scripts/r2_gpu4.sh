#!/usr/bin/env bash
# Round 2, single GPU: full GPU suite after the TC / LargeVis / dense-rows changes, c3 + c4 legs, re-order timing.
set -u
O=gpurun_out; mkdir -p $O
show() { python -c "
import json,sys
try:
    d=json.loads(open('$1').read()); r=d.get('roofline',{})
    print('$1', 'value', round(d['value'],2), d['unit'], 'ms/step', round(d['ms_per_step'],4), 'frac', round(r.get('frac',0),3), 'e2e', (d.get('e2e') or {}).get('seconds'), d.get('stage_ms') or d.get('reference_formulation_ms_per_iteration') or '', d.get('parity') if '${2:-}' else '', d['clocks']['sm_mhz'], d['clocks']['reasons'])
except Exception as e: print('$1 FAILED', e)
"; }
echo "== [1] pytest -m gpu"; timeout 1200 python -m pytest tests -q -m gpu 2>&1 | tail -12
echo "== [2] c3 full (100k x 256)"
timeout 600 python bench.py --config c3 > $O/r2_c3.json 2> $O/r2_c3.err; tail -2 $O/r2_c3.err; show $O/r2_c3.json p
echo "== [3] c4 at N=1 (LargeVis 10M x 64, row-local)"
timeout 900 python bench.py --config c4 --steps 10 > $O/r2_c4_n1.json 2> $O/r2_c4_n1.err; tail -2 $O/r2_c4_n1.err; show $O/r2_c4_n1.json
python -c "
import json; d=json.loads(open('gpurun_out/r2_c4_n1.json').read()); print({k: d.get(k) for k in ('affinity_seconds','union_graph_seconds','reference_formulation_ms_per_iteration','cpu_baseline')})"
echo "== [4] kNN by row order, 1M / 10M"
timeout 300 python scripts/knn_time.py 1000000 128 15 tree 2>&1 | tail -5
timeout 600 python scripts/knn_time.py 10000000 128 15 tree 2>&1 | tail -5
echo "== [5] e2e shuffled"
timeout 300 python bench.py --points 1000000 --order shuffled --steps 20 --no-cpu --no-parity > $O/r2_bench_1m_shuffled.json 2> $O/r2_bench_1m_shuffled.err; tail -2 $O/r2_bench_1m_shuffled.err
python -c "
import json; d=json.loads(open('gpurun_out/r2_bench_1m_shuffled.json').read()); print('shuffled 1M: value', round(d['value'],1), 'e2e s', round(d['e2e']['seconds'],3))"
echo "== [6] ncu launch list, 10M bench (kNN stage breakdown)"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file $O/r2_launches_10m.csv python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu --no-parity > $O/ncu_b.log 2>&1
python scripts/ncu_summary.py launches $O/r2_launches_10m.csv $O/r2_launches_10m_summary.txt; head -40 $O/r2_launches_10m_summary.txt
