#!/usr/bin/env bash
set -u
O=gpurun_out; mkdir -p $O
echo "== [1] LargeVis tests"; timeout 600 python -m pytest tests/test_gpu_parity.py -q -k "largevis or estimators_end_to_end" 2>&1 | tail -5
echo "== [2] e2e stage breakdown 1M"
timeout 300 python scripts/e2e_breakdown.py 1000000 generator 2>&1 | tail -2
timeout 300 python scripts/e2e_breakdown.py 1000000 shuffled 2>&1 | tail -2
echo "== [3] c4 at N=1 (row-local, MLP kernel)"
timeout 900 python bench.py --config c4 --steps 10 --no-cpu > $O/r2_c4_n1.json 2> $O/r2_c4_n1.err; tail -2 $O/r2_c4_n1.err
python -c "
import json; d=json.loads(open('gpurun_out/r2_c4_n1.json').read()); print('c4 n1: value', round(d['value'],2), 'ms/step', round(d['ms_per_step'],3), 'frac', round(d['roofline']['frac'],3), 'e2e', d['e2e']['seconds'], 'ref-form', d['reference_formulation_ms_per_iteration']['total'], 'nnz_union', d['config']['union_graph_nnz'], 'aff s', d['affinity_seconds'], d['union_graph_seconds'])"
