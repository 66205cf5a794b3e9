"""Oracle: fuzzy-union symmetrisation of the kNN graph (test infrastructure).

Restates ``torchdr/utils/sparse.py:7-206`` (``symmetrize_sparse`` with
``mode="sum_minus_prod"``: Q = P + P^T - P o P^T), whose output is a
-1-padded ELL pair ``(values[n, W], indices[n, W] int64)`` with columns in
ascending order (:118-140).  The CUDA engine keeps the same graph as CSR;
:func:`ell_to_csr` / :func:`csr_to_ell` convert between the two.
"""

import numpy as np
import torch


def symmetrize_ell(P, idx, mode="sum_minus_prod"):
    """``utils/sparse.py:170-206``."""
    n, k = P.shape
    rows = torch.arange(n).repeat_interleave(k)  # :28-35
    cols = idx.reshape(-1).long()
    vals = P.reshape(-1)
    # :62-81 — one sorted-unique over the keys of P and of P^T, two scatter-adds
    keys = torch.cat([rows * n + cols, cols * n + rows])
    uniq, inv = torch.unique(keys, sorted=True, return_inverse=True)
    m = uniq.numel()
    nv = vals.numel()
    from_p = torch.zeros(m, dtype=P.dtype).scatter_add_(0, inv[:nv], vals)
    from_pt = torch.zeros(m, dtype=P.dtype).scatter_add_(0, inv[nv:], vals)
    if mode == "sum":  # :161-162
        q = from_p + from_pt
    elif mode == "sum_minus_prod":  # :163-164
        q = from_p + from_pt - from_p * from_pt
    else:
        raise ValueError(f"Unsupported mode {mode!r}")
    ri = uniq // n  # :84-85
    ci = uniq % n
    # :118-140 — pack rows left-aligned, pad with (0, -1)
    deg = torch.bincount(ri, minlength=n)
    width = int(deg.max())
    out_v = torch.zeros((n, width), dtype=P.dtype)
    out_i = torch.full((n, width), -1, dtype=torch.long)
    start = torch.zeros(n + 1, dtype=torch.long)
    start[1:] = deg.cumsum(0)
    slot = torch.arange(m) - start[ri]
    out_v[ri, slot] = q
    out_i[ri, slot] = ci
    return out_v, out_i


def ell_to_csr(values, indices):
    """Drop the -1 padding: returns (rowptr int64[n+1], col int32[nnz], val[nnz])."""
    keep = indices >= 0
    deg = keep.sum(1)
    rowptr = torch.zeros(values.shape[0] + 1, dtype=torch.long)
    rowptr[1:] = deg.cumsum(0)
    return rowptr, indices[keep].int(), values[keep]


def csr_to_ell(rowptr, col, val, pad_val=0.0):
    rowptr = np.asarray(rowptr)
    n = rowptr.shape[0] - 1
    deg = rowptr[1:] - rowptr[:-1]
    width = int(deg.max()) if n else 0
    out_v = torch.full((n, width), pad_val, dtype=val.dtype)
    out_i = torch.full((n, width), -1, dtype=torch.long)
    r = torch.repeat_interleave(torch.arange(n), torch.as_tensor(deg))
    slot = torch.arange(col.shape[0]) - torch.as_tensor(rowptr[:-1])[r]
    out_v[r, slot] = val
    out_i[r, slot] = col.long()
    return out_v, out_i
