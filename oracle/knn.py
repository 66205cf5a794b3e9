"""Oracle: pairwise distances and exact kNN (test infrastructure, see package header).

Restates ``torchdr/distance/torch.py:21-125`` (torch backend of
``pairwise_distances``) and ``torchdr/utils/utils.py:173-216`` (``kmin``).
"""

import torch

_SELF_PENALTY = 1e12  # distance/torch.py:114 — added on the diagonal, not a skip


def _expanded_sq(Xq, Xdb, nq, ndb):
    # distance/torch.py:89-91: ||x||^2 (+) ||y||^2 - 2 X Y^T, all fp32, never clamped
    return nq.unsqueeze(-1) + ndb.unsqueeze(-2) - 2 * (Xq @ Xdb.transpose(-1, -2))


def pairwise_full(X, Y=None, metric="sqeuclidean", exclude_diag=False):
    """Full (n, m) matrix, ``distance/torch.py:63-116``."""
    if metric not in ("sqeuclidean", "euclidean"):
        raise ValueError(f"[TorchDR] ERROR : The '{metric}' distance is not supported.")
    same = Y is None or Y is X
    if same:
        Y = X
    nx = (X**2).sum(dim=-1)  # torch.py:81-82
    ny = nx if same else (Y**2).sum(dim=-1)
    C = _expanded_sq(X, Y, nx, ny)
    if metric == "euclidean":  # torch.py:92-95
        C = C.clamp(min=0).sqrt()
    if exclude_diag and same:  # torch.py:111-116
        pen = torch.zeros_like(C)
        r = torch.arange(C.shape[0])
        pen[r, r] = _SELF_PENALTY
        C = C + pen
    return C


def knn_dense(X, k, metric="sqeuclidean", exclude_diag=True):
    """(C[n,k], idx[n,k] int32) ascending; ``torch.py:119-122`` + ``utils.py:203-216``.

    ``k >= n`` returns the full matrix and ``None`` exactly like ``kmin``.
    """
    C = pairwise_full(X, None, metric, exclude_diag)
    if k >= C.shape[1]:
        return C, None
    vals, idx = C.topk(k=k, dim=1, largest=False)
    return vals, idx.int()


def knn_chunked(X, k, metric="sqeuclidean", exclude_diag=True, block=4096,
                q_start=0, q_end=None):
    """Block-by-block restatement for N beyond the dense N x N limit.

    Same arithmetic as :func:`knn_dense` applied to row blocks
    ``X[a:b]`` against the full database.  Not
    bit-identical to the dense sgemm (MKL blocking differs with M) but of the
    same accuracy class; validated against :func:`knn_dense` at N <= 20k.
    ``q_start/q_end`` restrict the query rows (the distributed chunk,
    ``distance/base.py:184-186``).
    """
    n = X.shape[0]
    q_end = n if q_end is None else q_end
    nx = (X**2).sum(dim=-1)
    out_v, out_i = [], []
    for a in range(q_start, q_end, block):
        b = min(a + block, q_end)
        C = _expanded_sq(X[a:b], X, nx[a:b], nx)
        if metric == "euclidean":
            C = C.clamp(min=0).sqrt()
        if exclude_diag:
            r = torch.arange(a, b)
            C[r - a, r] += _SELF_PENALTY
        v, i = C.topk(k=k, dim=1, largest=False)
        out_v.append(v)
        out_i.append(i.int())
    return torch.cat(out_v), torch.cat(out_i)


def knn_ambiguity(X, k, tau_rel=4e-6, block=2048, q_start=0, q_end=None):
    """Classify which kNN answers are decided at fp32 accuracy.

    Computes expanded-form squared distances in float64, takes the k+1
    smallest per row and measures the gaps between consecutive ranks.  A gap is
    *decided* when it exceeds ``tau_rel * (||x||^2 + ||y||^2)``: the fp32
    expanded form (``torch.py:89-91``) was measured at <= 9e-7 of that scale
    from the float64 value on the golden inputs, so two distances further apart
    than 4e-6 of it cannot be swapped by either the reference sgemm or an
    fp32 kernel of the same accuracy class (hard part 1).

    Returns ``(idx64[n,k], dist64[n,k], entry_ok[n,k] bool, set_ok[n] bool)``:
    ``entry_ok[i,r]`` — rank r of row i is separated from ranks r-1 and r+1;
    ``set_ok[i]`` — the k-th and (k+1)-th neighbours are separated, i.e. the
    neighbour SET of row i is decided.
    """
    n = X.shape[0]
    q_end = n if q_end is None else q_end
    Xd = X.double()
    nx = (Xd**2).sum(-1)
    idx_out, d_out, e_ok, s_ok = [], [], [], []
    kk = min(k + 1, n - 1)
    for a in range(q_start, q_end, block):
        b = min(a + block, q_end)
        D = nx[a:b, None] + nx[None, :] - 2.0 * (Xd[a:b] @ Xd.T)
        r = torch.arange(a, b)
        D[r - a, r] = float("inf")
        v, i = D.topk(kk, dim=1, largest=False)
        scale = nx[a:b, None] + nx[i]
        gap_ok = (v[:, 1:] - v[:, :-1]) > tau_rel * torch.maximum(scale[:, 1:], scale[:, :-1])
        if kk == k:  # no (k+1)-th neighbour exists: the last rank has no upper rival
            gap_ok = torch.cat([gap_ok, torch.ones(b - a, 1, dtype=torch.bool)], dim=1)
        left = torch.cat([torch.ones(b - a, 1, dtype=torch.bool), gap_ok[:, : k - 1]], dim=1)
        right = gap_ok[:, :k]
        e_ok.append(left & right)
        s_ok.append(gap_ok[:, k - 1])
        idx_out.append(i[:, :k])
        d_out.append(v[:, :k])
    return torch.cat(idx_out), torch.cat(d_out), torch.cat(e_ok), torch.cat(s_ok)


# --------------------------------------------------------------------------------------------------
# Restatement of the ENGINE's tile-pruning rule (torchdr_b200/csrc/knn_tc.cu, "pruned sweep").  The
# reference has no such step (distance/torch.py:81-122 always forms the full N x N matrix); this is
# test infrastructure that states, on the CPU, the rule the CUDA path must obey: a database tile may be
# skipped for a query tile only if no row of the query tile can have a neighbour in it.
def tile_prune_plan(X, k, tile=128, window=4, exclude_self=True):
    """Surviving-tile mask ``keep[n_tiles, n_tiles]`` (query tile x database tile) and the per-row bounds tau.

    tau_i: k-th smallest expanded-form distance of row i inside the +-``window`` tiles around its own
    tile (any upper bound on the k-th neighbour distance is admissible; the kernel uses group minima).
    keep[A, B] = box_distance^2(A, B) * 0.9999 <= max_{i in A} tau_i * 1.001 + 2e-5 * 2 * max ||x||^2,
    the same margins as ``tile_prune_kernel``."""
    X = X.float()
    n = X.shape[0]
    nt = (n + tile - 1) // tile
    nx = (X**2).sum(-1)
    lo = torch.stack([X[t * tile:(t + 1) * tile].min(0).values for t in range(nt)])
    hi = torch.stack([X[t * tile:(t + 1) * tile].max(0).values for t in range(nt)])
    tau = torch.empty(n)
    for t in range(nt):
        a, b = max(0, t - window), min(nt, t + window + 1)
        rows = slice(t * tile, min((t + 1) * tile, n))
        cols = slice(a * tile, min(b * tile, n))
        C = _expanded_sq(X[rows], X[cols], nx[rows], nx[cols])
        if exclude_self:
            r = torch.arange(C.shape[0])
            C[r, r + (rows.start - cols.start)] = float("inf")
        tau[rows] = C.kthvalue(k, dim=1).values
    bound = torch.stack([tau[t * tile:(t + 1) * tile].max() for t in range(nt)]) * 1.001 + 2e-5 * 2 * float(nx.max())
    keep = torch.zeros(nt, nt, dtype=torch.bool)
    for A in range(nt):
        gap = torch.maximum(torch.maximum(lo - hi[A], lo[A] - hi), torch.zeros(()))
        keep[A] = (gap * gap).sum(1) * 0.9999 <= bound[A]
    return keep, tau


def knn_with_tile_mask(X, k, keep, tile=128, exclude_self=True):
    """Exact kNN restricted to the database tiles ``keep`` allows (ascending, ties to the lower index)."""
    X = X.float()
    n = X.shape[0]
    nx = (X**2).sum(-1)
    nt = keep.shape[0]
    out_v, out_i = [], []
    for A in range(nt):
        rows = slice(A * tile, min((A + 1) * tile, n))
        C = _expanded_sq(X[rows], X, nx[rows], nx)
        if exclude_self:
            r = torch.arange(C.shape[0])
            C[r, r + rows.start] = float("inf")
        mask = keep[A].repeat_interleave(tile)[:n]
        C = torch.where(mask.unsqueeze(0), C, torch.full((), float("inf")))
        order = torch.argsort(C, dim=1, stable=True)[:, :k]
        out_v.append(C.gather(1, order))
        out_i.append(order.int())
    return torch.cat(out_v), torch.cat(out_i)
