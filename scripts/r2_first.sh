#!/usr/bin/env bash
# First GPU call of round 2: everything that was written at the end of round 1 without hardware.
#   1. the experimental robust pruned sweep (tdr_knn_set_prune(2)) against the full sweep, bit for bit
#   2. step kernel variants: default, L2 hints (5), persistent + prefetch (6), at 1 M and 10 M
#   3. shuffled row order: plain (full sweep) vs TDR_KNN_REORDER=1 (Voronoi-tree order + certified sweep)
set -u
O=gpurun_out; mkdir -p $O
echo "== [1] experimental kNN tests"; TDR_TEST_EXPERIMENTAL=1 timeout 300 python -m pytest tests/test_gpu_parity.py -x -q -s -k "robust or discard_nns or baseline_config_1 or generic_optimizers" 2>&1 | tail -12
echo "== [2] step kernel variants"
for cfg in 0 6 7 8 5; do
  TDR_STEP_CFG=$cfg timeout 200 python bench.py --steps 1000 --warmup 20 --no-e2e --no-cpu 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('1M cfg $cfg:', round(d['value'],1), 'it/s', round(d['ms_per_step'],4), 'ms')"
  TDR_STEP_CFG=$cfg timeout 300 python bench.py --points 10000000 --steps 200 --warmup 10 --no-e2e --no-cpu 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('10M cfg $cfg:', round(d['value'],1), 'it/s', round(d['ms_per_step'],4), 'ms', d['clocks']['sm_mhz'])"
done
for cfg in 6 7 8; do TDR_STEP_CFG=$cfg timeout 200 python -m pytest tests/test_gpu_parity.py -x -q -k "umap" 2>&1 | tail -2; done
echo "== [3] shuffled rows"
for re in 0 1; do
  TDR_KNN_REORDER=$re timeout 300 python bench.py --order shuffled --steps 200 --warmup 10 --no-cpu 2>$O/shuffled_$re.err | python -c "
import json,sys; d=json.loads(sys.stdin.read()); a=d['affinity_kernel']; print('shuffled reorder=$re: e2e', round(d['e2e']['seconds'],3), 's; bench kNN ms', round(a['ms'],1), 'swept', a['tile_pairs_swept'], 'of', a['tile_pairs_all'])"
done
