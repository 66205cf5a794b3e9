"""Dev: LargeVis / t-SNE iteration timing on one GPU (BASELINE configs 1/4 shapes, reduced N)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import torch
import torchdr_b200 as tb
from torchdr_b200 import ops
from helpers import clustered
dev = "cuda:0"
n, d = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000, 64
X = clustered(n, d).to(dev)
torch.cuda.synchronize(); t0 = time.perf_counter()
ea = tb.EntropicAffinity(perplexity=30, max_iter=100)
P, I = ea(X, log=False, return_indices=True)
torch.cuda.synchronize(); print(f"entropic affinity n={n} d={d} k=90: {time.perf_counter()-t0:.2f} s", flush=True)
Z = torch.randn(n, 2, device=dev) * 1e-4
grad = torch.zeros_like(Z); mom = torch.zeros_like(Z)
def lv_iter(t):
    grad.zero_(); ops.largevis_grad(Z, 0, n, P, I, grad, t, neg=None, n_neg=5, seed=1); ops.sgd_momentum(Z, mom, grad, 50.0, 0.8, t == 0)
for t in range(5): lv_iter(t)
torch.cuda.synchronize(); t0 = time.perf_counter()
for t in range(5, 55): lv_iter(t)
torch.cuda.synchronize(); dt = (time.perf_counter()-t0)/50
print(f"LargeVis iteration n={n}: {dt*1e3:.3f} ms ({1/dt:.0f} it/s), finite={bool(torch.isfinite(Z).all())}", flush=True)
nt = min(n, 100_000)
Zt = (torch.randn(nt, 2, device=dev) * 1e-4).contiguous(); g2 = torch.zeros_like(Zt); m2 = torch.zeros_like(Zt)
ws = ops.tsne_workspace(nt, dev); Pt, It = P[:nt].contiguous(), (I[:nt] % nt).contiguous()
def ts_iter(t):
    g2.zero_(); ops.tsne_grad(Zt, 0, nt, Pt, It, 12.0, 0, g2, ws); ops.tsne_grad(Zt, 0, nt, Pt, It, 12.0, 1, g2, ws); ops.sgd_momentum(Zt, m2, g2, 50.0, 0.5, t == 0)
for t in range(3): ts_iter(t)
torch.cuda.synchronize(); t0 = time.perf_counter()
for t in range(3, 13): ts_iter(t)
torch.cuda.synchronize(); dt = (time.perf_counter()-t0)/10
print(f"t-SNE iteration (dense N^2 repulsion) n={nt}: {dt*1e3:.2f} ms ({1/dt:.1f} it/s; {nt*nt/dt/1e9:.1f} G pairs/s)", flush=True)
