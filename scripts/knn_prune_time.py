"""Dev: fused tensor-core kNN at n x 128 with and without the pruned sweep (timing + launch list under ncu)."""
import sys, time
sys.path.insert(0, __import__("os").path.dirname(__import__("os").path.dirname(__import__("os").path.abspath(__file__))))
import torch
from bench import clustered
from torchdr_b200 import ops
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
modes = [int(m) for m in sys.argv[3].split(",")] if len(sys.argv) > 3 else [1]
X = clustered(n, 128, "cuda")
for on in modes:
    stats = torch.zeros(2, dtype=torch.int64, device="cuda")
    ops.knn_set_prune(bool(on), stats)
    for r in range(reps):
        stats.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize(); e0.record()
        out = ops.knn_umap_fused(X, X, 15, want_dist=False)
        e1.record(); torch.cuda.synchronize()
        print(f"n={n} prune={on} rep {r}: {e0.elapsed_time(e1):.2f} ms, swept/full = {stats.tolist()}", flush=True)
