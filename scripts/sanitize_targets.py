"""Small invocations of the hand-rolled synchronisation kernels for compute-sanitizer (memcheck / racecheck / synccheck):
the tcgen05 kNN kernel (mbarrier / TMEM / TMA pipeline; full, pruned and certified sweeps; fused rows), the persistent
UMAP loop (grid barrier on device words) and the per-iteration step kernel.  Shapes are tiny: the tools slow kernels
down by 10-100x.

  compute-sanitizer --tool memcheck python scripts/sanitize_targets.py
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch

from helpers import clustered
from torchdr_b200 import ops

dev = torch.device("cuda")
which = sys.argv[1] if len(sys.argv) > 1 else "all"
if which in ("all", "knn"):
    X = clustered(1500, 128).to(dev)
    ops.knn(X, X, 15, path="tc", prune="off")
    ops.knn_umap_fused(X, X, 15, path="tc", prune="off")
    ops.knn(X[:700].contiguous(), X, 33, path="tc")  # distinct query buffer, k at the shared-memory limit of d = 128
    X = clustered(9000, 64).to(dev)  # >= 64 tiles: phase A + box test + phase B
    ops.knn(X, X, 90, path="tc", prune="on")
    ops.knn_umap_fused(X, X, 15, path="tc", prune="certified")
    ops.knn(X[1000:5301], X, 15, q_row0=1000, path="tc", prune="certified")
    torch.cuda.synchronize()
    print("knn targets done")
if which in ("all", "step"):
    n, k = 6000, 15
    X = clustered(n, 32).to(dev)
    _, idx, P, _, _ = ops.knn_umap_fused(X, X, k, want_dist=False)
    rowptr, col, val = ops.symmetrize_csr(P, idx, 0, n)
    eps, _ = ops.umap_schedule(val, float(ops.max_value(val).item()), 50)
    rp, cc, ce, eons = ops.umap_compact(rowptr, col, eps)
    Z = (torch.randn(n, 2, device=dev) * 1e-4).contiguous()
    Zb = torch.empty_like(Z)
    sync = ops.RunSync(dev)
    ops.umap_run(Z, Zb, rp, cc, ce, eons, 0, [1.0, 0.9, 0.8, 0.7, 0.6], 1.58, 0.895, seed=3, sync=sync)
    ops.umap_step(Z, Zb, 0, n, rp, cc, ce, eons, 5, 1.58, 0.895, 0.5, neg=None, seed=3)
    ops.umap_step(Z, Zb, 0, n, rp, cc, ce, eons, 6, 1.58, 0.895, 0.5, neg=None, seed=3, precise=True)
    torch.cuda.synchronize()
    sync.check()
    print("step targets done")
