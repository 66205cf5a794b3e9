"""Timing of the kNN (+ sigma/rho) stage on one GPU: kernel path x prune mode x row order.

  python scripts/knn_time.py [n] [d] [k] [order: generator|shuffled|tree]
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch

from bench import clustered
from torchdr_b200 import ops
from torchdr_b200.reorder import index_locality, voronoi_tree_order

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
d = int(sys.argv[2]) if len(sys.argv) > 2 else 128
k = int(sys.argv[3]) if len(sys.argv) > 3 else 15
order = sys.argv[4] if len(sys.argv) > 4 else "generator"
dev = torch.device("cuda")
X = clustered(n, d, dev)
ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)


def timed(fn):
    torch.cuda.synchronize()
    ev0.record()
    out = fn()
    ev1.record()
    torch.cuda.synchronize()
    return out, ev0.elapsed_time(ev1)


if order != "generator":
    X = X[torch.randperm(n, device=dev, generator=torch.Generator(device=dev).manual_seed(7))].contiguous()
print(f"n={n} d={d} k={k} order={order}: index_locality {index_locality(X):.3f}")
if order == "tree":
    perm, ms = timed(lambda: voronoi_tree_order(X, generator=torch.Generator(device=dev).manual_seed(1)))
    X = X[perm].contiguous()
    print(f"  voronoi_tree_order: {ms:.1f} ms; index_locality after {index_locality(X):.3f}")
stats = torch.zeros(2, dtype=torch.int64, device=dev)
for prune in (["on", "certified"] + (["off"] if n <= 2_000_000 else [])):
    for rep in range(2):
        stats.zero_()
        if k <= 33:
            _, ms = timed(lambda: ops.knn_umap_fused(X, X, k, want_dist=False, prune=prune, sweep_stats=stats))
        else:
            _, ms = timed(lambda: ops.knn(X, X, k, prune=prune, sweep_stats=stats))
    sw, full = (int(v) for v in stats.tolist())
    print(f"  prune={prune:9s}: {ms:9.2f} ms   tile pairs swept {sw} of {full}" + (f" ({100.0 * sw / full:.2f} %)" if full else ""))
