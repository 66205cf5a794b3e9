"""Oracle: per-row bandwidth search affinities (test infrastructure).

Restates ``torchdr/affinity/knn_normalized.py:417-468`` (UMAPAffinity before
symmetrisation), ``torchdr/affinity/entropic.py:230-312`` (EntropicAffinity,
sparse log form) and ``entropic.py:51-115`` (Vladymyrov/Carreira-Perpinan
bracket).  Inputs are the kNN distance rows ``C[n, k]`` (ascending).
"""

import math

import torch

from .root_search import bisect_rows


def clamp_neighbor_param(value, n_samples):
    """``utils/validation.py:223-244`` tensor branch: long cast, clamp to [2, n-2]."""
    return int(min(max(int(value), 2), n_samples - 2))


def umap_affinity_rows(C, n_neighbors, max_iter=100):
    """``knn_normalized.py:445-468``.

    rho = row minimum; sigma solves  sum_j exp(-(C_ij - rho_i)/sigma_i) =
    log2(n_neighbors) with the marginal evaluated as exp(logsumexp(.)) (:452-454);
    returns ``(P, rho, sigma)`` with ``P = exp(-(C - rho)/sigma)`` (:464-465).
    """
    n = C.shape[0]
    rho = C.topk(1, dim=1, largest=False)[0].squeeze(1).contiguous()  # :445
    target = torch.log2(torch.tensor(n_neighbors, dtype=C.dtype))  # :448-450

    def gap(sig):
        lg = -(C - rho[:, None]) / sig[:, None]  # _log_P_UMAP, :40-42
        return lg.logsumexp(1).exp() - target

    sigma = bisect_rows(gap, n, 1.0, 1.0, max_iter=max_iter, dtype=C.dtype)  # :456-462
    P = (-(C - rho[:, None]) / sigma[:, None]).exp()
    return P, rho, sigma


def entropic_bounds(C, perplexity):
    """``entropic.py:51-115`` — returns ``(begin, end)`` *before* the +1e-6 of :287."""
    dtype = C.dtype
    tN = torch.tensor(C.shape[0], dtype=dtype)
    perp = torch.tensor(perplexity)  # long tensor in the reference (validation.py:229-244)
    cap = torch.minimum(torch.sqrt(2.0 * tN), perp)  # :73

    def p1_gap(x):  # :76-78
        return torch.log(cap) - 2.0 * (1.0 - x) * torch.log(tN / (2.0 * (1.0 - x)))

    lo = torch.tensor([0.75], dtype=dtype)
    hi = torch.tensor([1.0 - 1e-6], dtype=dtype)
    p1 = bisect_rows(p1_gap, 1, lo, hi, max_iter=1000, dtype=dtype).squeeze()  # :83-91

    dN = C.topk(1, dim=1, largest=True)[0].squeeze(1)  # :93
    d12 = C.topk(2, dim=1, largest=False)[0]  # :94
    d1, d2 = d12[:, 0], d12[:, 1]
    span = dN - d1
    step = d2 - d1
    lr = torch.log(tN / perp)  # :101
    beta_lo = torch.max(  # :102-105
        (tN * lr) / ((tN - 1) * span),
        torch.sqrt(lr / (dN.pow(2) - d1.pow(2))),
    )
    beta_hi = torch.log((tN - 1) * p1 / (1.0 - p1)) / step  # :106
    return 1 / beta_hi, 1 / beta_lo


def entropic_affinity_rows(C, perplexity, n_total=None, max_iter=100, use_bounds=True):
    """``entropic.py:272-310``.

    ``perplexity`` is the already-clamped integer (:257).  ``use_bounds=False``
    is the multi-GPU rule (:280-282): bracket starts from begin=end=1.
    Returns ``(log_P, eps, log_norm)`` where
    ``log_P = -C/eps - logsumexp(-C/eps) - log(n_total)``.
    """
    n = C.shape[0]
    n_total = n if n_total is None else n_total
    target = torch.log(torch.tensor(perplexity)) + 1  # :272 (long -> float32 log)

    def gap(eps):  # :274-277 with utils/utils.py:147-170 (log=True branch)
        lg = -C / eps[:, None]
        lg = lg - lg.logsumexp(1, keepdim=True)
        return -(lg.exp() * (lg - 1)).sum(1) - target

    if use_bounds:
        begin, end = entropic_bounds(C, perplexity)
        begin = begin + 1e-6  # :287
    else:
        begin, end = None, None
    eps = bisect_rows(gap, n, begin, end, max_iter=max_iter, dtype=C.dtype)  # :289-297
    lg = -C / eps[:, None]
    log_norm = lg.logsumexp(1, keepdim=True)  # :303
    lg = lg - log_norm
    lg = lg - torch.log(torch.tensor(n_total, dtype=C.dtype))  # :308-310
    return lg, eps, log_norm.squeeze(1)


def log2_target(k):
    return math.log2(k)
