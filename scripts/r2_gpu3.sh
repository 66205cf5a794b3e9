#!/usr/bin/env bash
# Round 2, single GPU: L2-residency experiment, bench legs c3/c4/c5 (small, then full c3/c4), compute-sanitizer.
set -u
O=gpurun_out; mkdir -p $O
show() { python -c "
import json,sys
try:
    d=json.loads(open('$1').read()); r=d.get('roofline',{})
    print('$1', 'value', round(d['value'],2), d['unit'], 'ms/step', round(d['ms_per_step'],4), 'frac', round(r.get('frac',0),3), 'e2e', (d.get('e2e') or {}).get('seconds'), d.get('stage_ms') or d.get('stage_ms_per_iteration') or '', d.get('parity') if '${2:-}' else '', d['clocks']['sm_mhz'], d['clocks']['reasons'])
except Exception as e: print('$1 FAILED', e)
"; }
echo "== [0] new TC shapes"
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -k "tensor_core or pruned_sweep or certified or pairwise_full or entropic_dense or golden" 2>&1 | tail -8
echo "== [1] L2 persistence, 10M"
for mb in 0 48 80; do
  TDR_L2_PERSIST_MB=$mb timeout 300 python bench.py --no-e2e --no-cpu --no-parity > $O/l2_$mb.json 2> $O/l2_$mb.err; grep "L2 persist" $O/l2_$mb.err; show $O/l2_$mb.json
done
echo "== [2] legs, small"
timeout 200 python bench.py --config c3 --points 20000 --steps 2 > $O/c3_small.json 2> $O/c3_small.err; tail -2 $O/c3_small.err; show $O/c3_small.json p
timeout 300 python bench.py --config c4 --points 200000 --steps 5 > $O/c4_small.json 2> $O/c4_small.err; tail -2 $O/c4_small.err; show $O/c4_small.json
timeout 300 python bench.py --config c5 --points 2000000 --steps 5 --no-cpu > $O/c5_small.json 2> $O/c5_small.err; tail -2 $O/c5_small.err; show $O/c5_small.json
echo "== [2b] entropic k=90 at 1M x 128: TC vs SIMT"
python - <<'PY'
import sys, torch
sys.path.insert(0, ".")
from bench import clustered
from torchdr_b200 import ops
X = clustered(1_000_000, 128, torch.device("cuda"))
ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
for path, prune, q in (("tc", "on", 1_000_000), ("tc", "off", 100_000), ("simt", None, 100_000)):
    for rep in range(2):
        torch.cuda.synchronize(); ev0.record()
        ops.knn(X[:q], X, 90, path=path, prune=prune)
        ev1.record(); torch.cuda.synchronize()
    print(f"k=90 1Mx128 path={path} prune={prune} queries={q}: {ev0.elapsed_time(ev1):.1f} ms", flush=True)
PY
echo "== [3] c3 full (100k x 256)"
timeout 600 python bench.py --config c3 > $O/r2_c3.json 2> $O/r2_c3.err; tail -2 $O/r2_c3.err; show $O/r2_c3.json p
echo "== [4] c4 full at N=1 (LargeVis 10M x 64)"
timeout 900 python bench.py --config c4 --steps 10 > $O/r2_c4_n1.json 2> $O/r2_c4_n1.err; tail -2 $O/r2_c4_n1.err; show $O/r2_c4_n1.json
echo "== [5] compute-sanitizer"
for tool in memcheck racecheck synccheck; do
  timeout 900 compute-sanitizer --tool $tool --error-exitcode 9 python scripts/sanitize_targets.py > $O/sanitizer_$tool.log 2>&1; echo "$tool exit $?"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|targets done|Error|hazard" $O/sanitizer_$tool.log | head -12
done
