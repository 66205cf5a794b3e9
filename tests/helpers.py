"""Shared test helpers: golden loading, seeded inputs, negative tables."""

import os

import numpy as np
import torch

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def golden(name):
    z = np.load(os.path.join(GOLDEN, name + ".npz"))
    return {k: z[k] for k in z.files}


def t(a):
    return torch.from_numpy(np.ascontiguousarray(a))


def neg_seed(seed, step):  # same rule as tests/golden/make_golden.py
    return seed * 1000003 + step


def raw_negatives(seed, step, n_total, n_rows, n_neg):
    """The draw the reference made at `step` (NE base.py:629): randint(0, N-1)."""
    g = torch.Generator().manual_seed(neg_seed(seed, step))
    return torch.randint(0, n_total - 1, (n_rows, n_neg), generator=g)


def negative_table(seed, step, n_total, n_neg, row0=0, n_rows=None):
    """Adjusted negatives (self index skipped), int64 [n_rows, n_neg]."""
    n_rows = n_total if n_rows is None else n_rows
    raw = raw_negatives(seed, step, n_total, n_rows, n_neg)
    me = torch.arange(row0, row0 + n_rows).unsqueeze(1)
    return raw + (raw >= me).long()


def blobs(n, d, centers, seed, spread=1.0, scale=6.0):
    g = torch.Generator().manual_seed(seed)
    c = torch.randn(centers, d, generator=g) * scale
    lab = torch.randint(0, centers, (n,), generator=g)
    return (c[lab] + torch.randn(n, d, generator=g) * spread).float().contiguous()


def clustered(n, d, seed=42):
    """The reference benchmark's generator (benchmarks/faiss/run_benchmark.py:127-146)."""
    g = torch.Generator().manual_seed(seed)
    nc = max(1, min(1000, n // 100))
    centers = torch.randn(nc, d, generator=g) * 10
    per = n // nc
    parts = []
    for i in range(nc):
        m = per if i < nc - 1 else n - per * (nc - 1)
        parts.append(centers[i] + torch.randn(m, d, generator=g) * 0.5)
    return torch.cat(parts).float().contiguous()


def rel_fro(a, b):
    a = torch.as_tensor(a, dtype=torch.float64)
    b = torch.as_tensor(b, dtype=torch.float64)
    return float((a - b).norm() / b.norm().clamp_min(1e-300))
