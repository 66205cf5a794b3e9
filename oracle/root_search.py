"""Oracle: batched bracket-and-bisect root finder (test infrastructure).

Restates ``torchdr/utils/root_search.py:17-77`` (``binary_search``) and
``:147-198`` (``init_bounds``).  Rows are independent, so the CUDA kernels run
this loop per row; this batched form keeps the reference's global stopping
rule ("stop when no row is active") for bit-level comparisons.
"""

import torch

_TOL = 1e-6  # root_search.py:13


def _as_vec(v, n, dtype):
    if isinstance(v, torch.Tensor):
        v = v.to(dtype=dtype)
        if v.shape != (n,):
            raise ValueError(f"bound tensor must have shape ({n},), got {v.shape}")
        return v.clone()
    return torch.full((n,), 1.0 if v is None else float(v), dtype=dtype)


def bracket_rows(f, n, begin=1.0, end=1.0, max_iter=100, dtype=torch.float32):
    """``init_bounds`` (root_search.py:147-198)."""
    lo = _as_vec(begin, n, dtype)
    hi = _as_vec(end, n, dtype)
    # phase 1 (root_search.py:176-185): halve lo while f(lo) > 0, dragging hi down
    for _ in range(max_iter):
        pos = f(lo) > 0
        if not bool(pos.any()):
            break
        hi = torch.where(pos, torch.minimum(hi, lo), hi)
        lo = torch.where(pos, lo * 0.5, lo)
    # phase 2 (root_search.py:187-196): double hi while f(hi) < 0, dragging lo up
    for _ in range(max_iter):
        neg = f(hi) < 0
        if not bool(neg.any()):
            break
        lo = torch.where(neg, torch.maximum(lo, hi), lo)
        hi = torch.where(neg, hi * 2.0, hi)
    return lo, hi


def bisect_rows(f, n, begin=1.0, end=1.0, max_iter=100, dtype=torch.float32,
                return_evals=False):
    """``binary_search`` (root_search.py:17-77).  Returns the last midpoint."""
    tol = torch.tensor(_TOL).to(dtype)
    lo, hi = bracket_rows(f, n, begin, end, max_iter, dtype)
    f_lo = f(lo)
    mid = (lo + hi) * 0.5
    f_mid = f(mid)
    evals = 0
    for _ in range(max_iter):
        live = f_mid.abs() >= tol  # root_search.py:60
        if not bool(live.any()):
            break
        same = f_mid * f_lo > 0  # root_search.py:64
        up = live & same
        down = live & ~same
        lo = torch.where(up, mid, lo)
        f_lo = torch.where(up, f_mid, f_lo)
        hi = torch.where(down, mid, hi)
        mid = (lo + hi) * 0.5
        f_mid = f(mid)
        evals += 1
    return (mid, evals) if return_evals else mid
