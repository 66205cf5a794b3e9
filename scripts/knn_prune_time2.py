"""Dev: pruned kNN at 1 M x 128 — fused vs plain, standalone affinity kernel."""
import sys, time
sys.path.insert(0, __import__("os").path.dirname(__import__("os").path.dirname(__import__("os").path.abspath(__file__))))
import torch
from bench import clustered
from torchdr_b200 import ops
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
X = clustered(n, 128, "cuda")
def timed(fn, reps=3):
    out = None
    for r in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize(); e0.record(); out = fn(); e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
    return out, ms
(_, I, P, rho, sig), ms = timed(lambda: ops.knn_umap_fused(X, X, 15, want_dist=False)); print("fused pruned", ms)
(C, I2), ms = timed(lambda: ops.knn(X, X, 15)); print("plain pruned", ms)
(P2, rho2, sig2), ms = timed(lambda: ops.umap_affinity_rows(C)); print("standalone affinity rows", ms)
print("equal:", torch.equal(I, I2), torch.equal(P, P2), torch.equal(sig, sig2))
