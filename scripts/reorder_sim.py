"""Design study (CPU, numpy) for round 2: exact pruned kNN on inputs WITHOUT index locality.

The pruned sweep of csrc/knn_tc.cu needs rows that are close in space to be close in index.  This script
simulates the two additions that make it independent of the input order, and counts the tile pairs they sweep:

 1. re-ordering: the rows are sorted by the leaves of a tree.  Two trees are compared: a balanced random-projection
    tree (every node: project on the difference of two random member points, cut at the median; every tile is one
    leaf) and an unbalanced Voronoi tree (every node: nearest of <= 16 centres sampled from the node, recurse until
    <= 128 rows).  Measured at 50 k x 128, k = 15, shuffled input (tile pairs swept, first + second sweep):
        1000-point clusters: original order 2.5 %, shuffled 100 %, RP tree 60.9 + 2.1 %, Voronoi tree 3.2 %
         100-point clusters: original order 2.3 %, shuffled 100 %, Voronoi tree 2.9 % (leaves <= 512: 69.9 %)
    Median cuts slice clusters into fragments that end up in many tiles; the Voronoi tree keeps them together;
 2. guess - sweep - certify: a tile's pruning threshold T is taken from the rows whose phase-A bound tau_i is not
    an outlier (tau_i <= c x the tile's median); after the sweep, row i is certified exact iff its k-th distance
    found, tau'_i, satisfies tau'_i * 1.001 + margin <= T (every tile that could hold a closer point was swept,
    because box-to-box <= point-to-box distance); tiles with uncertified rows are swept again with
    T2 = max tau'_i, which is tight by then.

    python scripts/reorder_sim.py [n] [d] [k] [clustered|bigclusters|uniform] [c]
"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bench import clustered  # noqa: E402

BM = 128


def rp_tree_order(X, leaf=BM, seed=0):
    """Permutation that sorts rows by the leaves of a balanced random-projection tree."""
    rng = np.random.default_rng(seed)
    n = X.shape[0]
    perm = np.arange(n)
    segs = [(0, n)]
    while segs:
        nxt = []
        for a, b in segs:
            m = b - a
            if m <= leaf:
                continue
            idx = perm[a:b]
            p, q = idx[rng.integers(m)], idx[rng.integers(m)]
            direction = (X[p] - X[q]).astype(np.float64)
            if not direction.any():
                direction = rng.standard_normal(X.shape[1])
            proj = X[idx].astype(np.float64) @ direction
            half = ((m // 2 + leaf - 1) // leaf) * leaf  # cut on a tile boundary
            order = np.argpartition(proj, half - 1) if half < m else np.arange(m)
            left = order[:half]
            right = order[half:]
            perm[a:a + half] = idx[left]
            perm[a + half:b] = idx[right]
            nxt += [(a, a + half), (a + half, b)]
        segs = nxt
    return perm


def voronoi_tree_order(X, branch=16, leaf=BM, seed=0, lloyd=1):
    """Permutation that sorts rows by the leaves (depth-first) of an unbalanced Voronoi tree: every node samples
    up to `branch` of its own points as centres, runs `lloyd` Lloyd iterations and hands each point to its nearest
    centre; nodes of at most `leaf` rows are leaves.  Re-sampling inside every node is what defeats the hub effect
    of a flat random-centre quantiser in high dimension (a few centres collect most orphan clusters)."""
    rng = np.random.default_rng(seed)
    out = []

    def rec(idx, depth):
        m = len(idx)
        if m <= leaf or depth > 16:
            out.append(idx)
            return
        b = min(branch, max(2, m // leaf))
        P = X[idx].astype(np.float64)
        cen = P[rng.choice(m, b, replace=False)]
        for it in range(lloyd + 1):
            a = ((P ** 2).sum(1)[:, None] + (cen ** 2).sum(1)[None, :] - 2 * P @ cen.T).argmin(1)
            if it < lloyd:
                for j in range(b):
                    if (a == j).any():
                        cen[j] = P[a == j].mean(0)
        if np.bincount(a, minlength=b).max() == m:  # duplicates: cannot split
            out.append(idx)
            return
        for j in range(b):
            sub = idx[a == j]
            if len(sub):
                rec(sub, depth + 1)

    rec(np.arange(X.shape[0]), 0)
    return np.concatenate(out)


def boxes(X):
    nt = (X.shape[0] + BM - 1) // BM
    lo = np.stack([X[t * BM:(t + 1) * BM].min(0) for t in range(nt)])
    hi = np.stack([X[t * BM:(t + 1) * BM].max(0) for t in range(nt)])
    return lo, hi


def box_d2(lo, hi, a):
    gap = np.maximum(0, np.maximum(lo - hi[a][None, :], lo[a][None, :] - hi))
    return (gap * gap).sum(1)


def study(X, k, c_out=4.0, W=4):
    n, d = X.shape
    nt = (n + BM - 1) // BM
    Xd = X.astype(np.float64)
    nn = (Xd ** 2).sum(1)
    margin = 2e-5 * 2 * nn.max()
    lo, hi = boxes(X)
    swept1 = swept2 = 0
    failing_rows = failing_tiles = 0
    wrong = 0
    for t in range(nt):
        rows = slice(t * BM, min((t + 1) * BM, n))
        q = Xd[rows]
        D = (q ** 2).sum(1)[:, None] + nn[None, :] - 2 * q @ Xd.T
        r = np.arange(q.shape[0])
        D[r, t * BM + r] = np.inf
        # phase A: exact k-th inside the +-W window
        a, b = max(0, t - W) * BM, min(nt, t + W + 1) * BM
        tau = np.partition(D[:, a:b], k - 1, axis=1)[:, k - 1]
        med = np.median(tau)
        T = tau[tau <= c_out * med].max() * 1.001 + margin
        keep = box_d2(lo, hi, t) * 0.9999 <= T
        swept1 += int(keep.sum())
        mask = np.repeat(keep, BM)[:n]
        Dm = np.where(mask[None, :], D, np.inf)
        tau1 = np.partition(Dm, k - 1, axis=1)[:, k - 1]
        bad = tau1 * 1.001 + margin > T
        true_kth = np.partition(D, k - 1, axis=1)[:, k - 1]
        if bad.any():
            failing_rows += int(bad.sum())
            failing_tiles += 1
            T2 = tau1[bad].max() * 1.001 + margin
            keep2 = box_d2(lo, hi, t) * 0.9999 <= max(T, T2)
            swept2 += int(keep2.sum())
            mask2 = np.repeat(keep2, BM)[:n]
            tau2 = np.partition(np.where(mask2[None, :], D, np.inf), k - 1, axis=1)[:, k - 1]
            wrong += int((tau2 != true_kth).sum())
        else:
            wrong += int((tau1 != true_kth).sum())
    full = nt * nt
    print(f"  tiles {nt}: first sweep {swept1} pairs ({100 * swept1 / full:.2f} %), {failing_rows} uncertified rows in "
          f"{failing_tiles} tiles, second sweep {swept2} pairs ({100 * swept2 / full:.2f} %), rows with a wrong k-th "
          f"distance after certification: {wrong}")


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 50_000
    d = int(sys.argv[2]) if len(sys.argv) > 2 else 128
    k = int(sys.argv[3]) if len(sys.argv) > 3 else 15
    kind = sys.argv[4] if len(sys.argv) > 4 else "bigclusters"
    c_out = float(sys.argv[5]) if len(sys.argv) > 5 else 4.0
    if kind == "uniform":
        X = torch.randn(n, d, generator=torch.Generator().manual_seed(0)).numpy()
    elif kind == "bigclusters":  # cluster size of the 1 M benchmark (1000 points) at a smaller n
        g = torch.Generator().manual_seed(42)
        centers = torch.randn(n // 1000, d, generator=g) * 10
        X = (centers.repeat_interleave(1000, 0) + torch.randn(n, d, generator=g) * 0.5).numpy()
    else:
        X = clustered(n, d, "cpu").numpy()
    print(f"{kind} n={n} d={d} k={k}")
    print(" original order:")
    study(X, k, c_out)
    Xs = X[np.random.default_rng(0).permutation(n)]
    print(" shuffled:")
    study(Xs, k, c_out)
    print(" shuffled, then re-ordered by a balanced random-projection tree:")
    study(Xs[rp_tree_order(Xs)], k, c_out)
    print(" shuffled, then re-ordered by an unbalanced Voronoi tree (branching 16, leaves <= 128 rows, 1 Lloyd iteration):")
    study(Xs[voronoi_tree_order(Xs)], k, c_out)


if __name__ == "__main__":
    main()
