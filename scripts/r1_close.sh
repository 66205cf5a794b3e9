#!/usr/bin/env bash
# Final verification of the round: GPU suite, default bench, ncu launch list of the bench command.
set -u
O=gpurun_out; mkdir -p $O
timeout 300 python -m pytest tests -m gpu -q > $O/pytest_gpu.log 2>&1; tail -3 $O/pytest_gpu.log
timeout 400 python bench.py > $O/r1_bench.json 2> $O/r1_bench.err; tail -c 200 $O/r1_bench.err
python -c "
import json; d=json.load(open('$O/r1_bench.json')); print('value', d['value'], 'e2e', d['e2e']['value'], d['e2e']['seconds'], 'knn ms', d['affinity_kernel']['ms'], d['affinity_kernel'].get('full_sweep_ms'), 'frac', d['roofline']['frac'], d['clocks'])"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 2500 --csv --log-file $O/r1_launches.csv \
  python bench.py --steps 50 --warmup 5 --no-cpu > $O/launches.log 2>&1
python scripts/ncu_summary.py launches $O/r1_launches.csv $O/r1_launches_summary.txt | head -12
