timeout 300 python -m pytest tests/test_gpu_parity.py -x -q -s -k "pruned" 2>&1 | tail -15
timeout 200 python - <<'PY'
import torch, time, sys
sys.path.insert(0, '.')
from bench import clustered
from torchdr_b200 import ops
for n in (1_000_000,):
    X = clustered(n, 128, "cuda")
    for on in (0, 1, 1):
        stats = torch.zeros(2, dtype=torch.int64, device="cuda")
        ops.knn_set_prune(bool(on), stats)
        torch.cuda.synchronize(); t0 = time.perf_counter()
        out = ops.knn_umap_fused(X, X, 15, want_dist=False)
        torch.cuda.synchronize(); dt = time.perf_counter() - t0
        print(f"n={n} prune={on}: {dt*1e3:.1f} ms, swept/full = {stats.tolist()}", flush=True)
        if on == 0: ref = out
        else: print("  identical:", all(torch.equal(a, b) for a, b in zip(ref[1:], out[1:])))
PY
