"""Design check (CPU, numpy) of the tile-pruned exact kNN sweep of csrc/knn_tc.cu.

Restates the pruning plan on the host: per 128-row tile a ball (centre, radius); phase A = exact kNN of
every query row inside a window of +-W tiles -> an upper bound tau_i on the row's k-th neighbour distance;
a database tile B is swept for query tile A only if  (|cA - cB| - rA - rB)^2 <= max_i tau_i (+ margin).
Prints the fraction of (query tile, database tile) pairs that survive and checks on a subsample of rows that
the kNN restricted to surviving tiles equals the brute-force kNN.

    python scripts/prune_sim.py [n] [d] [k] [clustered|uniform|shuffled]
"""
import sys

import numpy as np
import torch

sys.path.insert(0, __import__("os").path.dirname(__import__("os").path.dirname(__import__("os").path.abspath(__file__))))
from bench import clustered  # noqa: E402

BM = 128


def plan(X, k, W=8, margin_rel=1e-3, margin_abs_norm=1e-4):
    n, d = X.shape
    nt = (n + BM - 1) // BM
    cen = np.zeros((nt, d), np.float32)
    rad = np.zeros(nt, np.float32)
    for t in range(nt):
        blk = X[t * BM:(t + 1) * BM]
        cen[t] = blk.mean(0)
        rad[t] = np.sqrt(((blk - cen[t]) ** 2).sum(1)).max() * (1 + 1e-4)
    norms = (X.astype(np.float64) ** 2).sum(1)
    T = np.zeros(nt, np.float32)
    for t in range(nt):
        lo, hi = max(0, t - W), min(nt, t + W + 1)
        q = X[t * BM:(t + 1) * BM].astype(np.float64)
        db = X[lo * BM:hi * BM].astype(np.float64)
        D = (q ** 2).sum(1)[:, None] + (db ** 2).sum(1)[None, :] - 2 * q @ db.T
        rows = np.arange(q.shape[0])
        D[rows, t * BM - lo * BM + rows] = np.inf  # exclude self
        T[t] = np.partition(D, k - 1, axis=1)[:, k - 1].max()
    Tpad = T * (1 + margin_rel) + margin_abs_norm * 2 * norms.max()
    cd = np.sqrt(np.maximum(((cen ** 2).sum(1)[:, None] + (cen ** 2).sum(1)[None, :] - 2 * cen @ cen.T), 0)) * (1 - 1e-4)
    lb = cd - rad[:, None] - rad[None, :]
    survive_ball = (lb <= 0) | (lb * lb <= Tpad[:, None])
    # axis-aligned boxes: sum_d max(0, lo_B - hi_A, lo_A - hi_B)^2
    lo = np.stack([X[t * BM:(t + 1) * BM].min(0) for t in range(nt)])
    hi = np.stack([X[t * BM:(t + 1) * BM].max(0) for t in range(nt)])
    survive_box = np.zeros((nt, nt), bool)
    for a in range(nt):
        gap = np.maximum(0, np.maximum(lo - hi[a][None, :], lo[a][None, :] - hi))
        survive_box[a] = (gap * gap).sum(1) * (1 - 1e-4) <= Tpad[a]
    print(f"  ball: {survive_ball.sum(1).mean():.1f} tiles/query tile, box: {survive_box.sum(1).mean():.1f}, both: "
          f"{(survive_ball & survive_box).sum(1).mean():.1f} (max {(survive_ball & survive_box).sum(1).max()})")
    return survive_ball & survive_box, T


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 100_000
    d = int(sys.argv[2]) if len(sys.argv) > 2 else 128
    k = int(sys.argv[3]) if len(sys.argv) > 3 else 15
    kind = sys.argv[4] if len(sys.argv) > 4 else "clustered"
    if kind == "uniform":
        X = torch.randn(n, d, generator=torch.Generator().manual_seed(0)).numpy()
    else:
        X = clustered(n, d, "cpu").numpy()
        if kind.startswith("big"):  # cluster size of the 1 M benchmark (1000 points) at a smaller n
            g = torch.Generator().manual_seed(42)
            nc = n // 1000
            centers = torch.randn(nc, d, generator=g) * 10
            X = (centers.repeat_interleave(1000, 0) + torch.randn(n, d, generator=g) * 0.5).numpy()
        if kind.endswith("shuffled") or kind.endswith("reordered"):
            X = X[np.random.default_rng(0).permutation(n)]
        if kind.endswith("reordered"):  # coarse-quantiser pass: nearest of C sampled points, stable sort by cell
            rng = np.random.default_rng(1)
            C = max(64, n // int(__import__("os").environ.get("CELL", "256")))
            cen = X[rng.choice(n, C, replace=False)].astype(np.float64)
            cn = (cen ** 2).sum(1)
            cell = np.empty(n, np.int64)
            for a in range(0, n, 8192):
                blk = X[a:a + 8192].astype(np.float64)
                cell[a:a + 8192] = ((blk ** 2).sum(1)[:, None] + cn[None, :] - 2 * blk @ cen.T).argmin(1)
            X = X[np.argsort(cell, kind="stable")]
    survive, T = plan(X, k)
    nt = survive.shape[0]
    per = survive.sum(1)
    print(f"{kind} n={n} d={d} k={k}: tiles {nt}, surviving tiles per query tile: mean {per.mean():.1f} "
          f"max {per.max()} ({100 * survive.mean():.2f}% of all pairs)")
    # exactness on a subsample of query tiles
    rng = np.random.default_rng(1)
    bad = 0
    Xd = X.astype(np.float64)
    nn = (Xd ** 2).sum(1)
    for t in rng.choice(nt, size=min(nt, 24), replace=False):
        q = Xd[t * BM:(t + 1) * BM]
        D = (q ** 2).sum(1)[:, None] + nn[None, :] - 2 * q @ Xd.T
        rows = np.arange(q.shape[0])
        D[rows, t * BM + rows] = np.inf
        full = np.argsort(D, axis=1, kind="stable")[:, :k]
        mask = np.repeat(survive[t], BM)[:n]
        Dm = np.where(mask[None, :], D, np.inf)
        pruned = np.argsort(Dm, axis=1, kind="stable")[:, :k]
        bad += int((full != pruned).sum())
    print("index mismatches between pruned and brute-force kNN on sampled tiles:", bad)


if __name__ == "__main__":
    main()
