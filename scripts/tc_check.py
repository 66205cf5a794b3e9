"""Dev check (GPU): tensor-core kNN path against the SIMT path and fp64, plus timings."""
import ctypes, sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import torch
from torchdr_b200 import ops, _lib
from helpers import blobs, clustered
import oracle

lib = _lib.load()
dev = "cuda:0"

def run(path, X, k, fused=False):
    lib.tdr_knn_set_path(path)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    if fused:
        out = ops.knn_umap_fused(X, X, k)
    else:
        out = ops.knn(X, X, k)
    torch.cuda.synchronize()
    return out, time.perf_counter() - t0

for (n, d, k) in [(300, 16, 15), (1000, 50, 7), (1500, 128, 15), (5000, 64, 30), (4096, 96, 15)]:
    X = blobs(n, d, 6, n + d)
    Xd = X.to(dev)
    (C1, I1), _ = run(1, Xd, k)
    (C2, I2), _ = run(2, Xd, k)
    idx64, d64, ok, set_ok = oracle.knn_ambiguity(X, k)
    scale = float((X ** 2).sum(1).max()) * 2
    e1 = float((C1.cpu().double() - d64).abs().max()) / scale
    e2 = float((C2.cpu().double() - d64).abs().max()) / scale
    m_ok = bool(torch.equal(I2.cpu().long()[ok], idx64[ok]))
    print(f"n={n} d={d} k={k}: simt err/scale {e1:.2e}  tc err/scale {e2:.2e}  tc==fp64 on decided: {m_ok}  "
          f"tc==simt idx frac {float((I1 == I2).float().mean()):.4f} decided frac {float(ok.float().mean()):.3f}", flush=True)

for n in (100_000, 1_000_000):
    X = clustered(n, 128).to(dev)
    for path in (2, 1):
        if path == 1 and n > 200_000:
            continue
        out, dt = run(path, X, 15, fused=True)
        out, dt = run(path, X, 15, fused=True)
        print(f"n={n} path={path} fused kNN+sigma: {dt*1e3:.1f} ms  ({2*n*n*128/dt/1e12:.1f} TFLOP/s fp32-equivalent)", flush=True)
    if n <= 200_000:
        (d1, i1, P1, r1, s1), _ = run(1, X, 15, fused=True)
        (d2, i2, P2, r2, s2), _ = run(2, X, 15, fused=True)
        print("  idx agree frac", float((i1 == i2).float().mean()), "max |dist diff|", float((d1 - d2).abs().max()),
              "sigma rel", float(((s1 - s2).abs() / s1).max()))
