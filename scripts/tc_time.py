"""Dev: time the fused tensor-core kNN at n (TDR_TC_DEBUG variants are timing experiments, results invalid)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import torch
from torchdr_b200 import ops, _lib
from helpers import clustered
n = int(sys.argv[1]) if len(sys.argv) > 1 else 200_000
_lib.load().tdr_knn_set_path(2)
X = clustered(n, 128).to("cuda:0")
for _ in range(2):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    ops.knn_umap_fused(X, X, 15)
    torch.cuda.synchronize(); dt = time.perf_counter() - t0
tiles = ((n + 127) // 128) ** 2
print(f"debug={os.environ.get('TDR_TC_DEBUG','0')} n={n}: {dt*1e3:.1f} ms, {dt / (tiles / 148) * 1e6:.2f} us per tile per SM", flush=True)
