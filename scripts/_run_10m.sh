timeout 200 python -m pytest tests/test_gpu_parity.py -x -q -s -k "pruned or knn_cross or golden" 2>&1 | tail -7
python scripts/knn_prune_time2.py 1000000
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/prune_launches3.csv python scripts/knn_prune_time.py 1000000 1 1 > /dev/null 2>&1; python scripts/ncu_summary.py launches gpurun_out/prune_launches3.csv gpurun_out/prune_launches3.txt | grep -E "tdr::|captured"
timeout 300 python bench.py --points 10000000 --steps 300 --warmup 10 --no-cpu > gpurun_out/r1_bench_10m.json 2> gpurun_out/r1_bench_10m.err; tail -c 300 gpurun_out/r1_bench_10m.err
python -c "
import json; d=json.load(open('gpurun_out/r1_bench_10m.json')); print(d['value'], d['e2e']['value'], d['e2e']['seconds'], d['affinity_kernel']['ms'], d['affinity_kernel']['tile_pairs_swept'], d['affinity_kernel']['tile_pairs_all'])"
