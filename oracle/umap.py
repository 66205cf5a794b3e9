"""Oracle: UMAP edge schedule, closed-form gradients and SGD loop (test infrastructure).

Restates ``torchdr/neighbor_embedding/umap.py:19-36`` (a/b fit), ``:215-234``
(edge schedule), ``:236-292`` (gradients), ``neighbor_embedding/base.py:617-636``
(negative sampling adjustment), ``:175-182`` + ``affinity_matcher.py:594-642``
(SGD + LinearLR 1 -> 0) and the loop body ``affinity_matcher.py:308-430``.
The graph is the reference's -1-padded ELL ``(values, indices)``.
"""

import numpy as np
import torch


def find_ab(spread=1.0, min_dist=0.1):
    """``umap.py:19-36`` — scipy curve fit of 1/(1+a x^(2b)) to the offset exponential."""
    from scipy.optimize import curve_fit

    def curve(x, a, b):
        return 1.0 / (1.0 + a * x ** (2 * b))

    xs = np.linspace(0, spread * 3, 300)
    ys = np.where(xs < min_dist, 1.0, np.exp(-(xs - min_dist) / spread))
    (a, b), _ = curve_fit(curve, xs, ys)
    return float(a), float(b)


def umap_edge_schedule(values, max_iter):
    """``umap.py:215-234``: returns (epochs_per_sample, epoch_of_next_sample)."""
    top = values.max()
    weak = values <= top / max_iter  # :219-221
    per = (values + 1e-3).reciprocal() * top  # :227  add_(1e-3).reciprocal_().mul_(A_max)
    per = per.masked_fill(weak, float("inf"))  # :228-230
    return per, per.clone()


def adjust_negatives(raw, self_idx):
    """``neighbor_embedding/base.py:629-636``: raw in [0, N-2] -> skip the row's own index."""
    return raw + (raw >= self_idx.unsqueeze(1)).long()


def linear_lr_sequence(lr0, max_iter, n_steps, start=1.0, end=0.0, total=None):
    """fp32 learning rates actually applied at steps 0..n_steps-1.

    Uses the same ``torch.optim`` objects as the reference call sites
    (``NE base.py:312-343`` SGD with a float lr; ``NE base.py:176-182`` LinearLR
    with tensor factors), on a dummy parameter.
    """
    p = torch.nn.Parameter(torch.zeros(1))
    opt = torch.optim.SGD([p], lr=lr0)  # NE base.py:343 passes the python float
    sch = torch.optim.lr_scheduler.LinearLR(
        opt,
        start_factor=torch.tensor(start),
        end_factor=torch.tensor(end),
        total_iters=max_iter if total is None else total,
    )
    out = []
    for _ in range(n_steps):
        out.append(float(opt.param_groups[0]["lr"]))
        p.grad = torch.zeros(1)
        opt.step()
        sch.step()
    return np.asarray(out, dtype=np.float32)


def _sq_to_rows(Zq, Zk):
    # distance/base.py:384-385 — exact-difference form
    return torch.sum((Zq.unsqueeze(1) - Zk) ** 2, dim=-1)


def umap_step(Z, idx, per, nxt, neg, n_iter, a, b, chunk_start=0, chunk_size=None,
              negative_sample_rate=5, eps=1e-3, lam=1.0, repulsion=1.0):
    """One call of ``_compute_gradients`` (NE base.py:235-242) for one row chunk.

    ``idx/per/nxt`` are the chunk's ELL rows; ``nxt`` is updated in place
    (umap.py:251-255).  Returns the chunk gradient ``[chunk_size, q]``.
    """
    n_local = idx.shape[0] if chunk_size is None else chunk_size
    rows = torch.arange(chunk_start, chunk_start + n_local)
    Zq = Z[rows]
    # --- attraction, umap.py:236-264
    Zk = Z[idx.long()]
    D = _sq_to_rows(Zq, Zk)
    pos = D > 0
    den = 1 + a * D**b
    D = D.pow(b - 1)
    D = D.mul(2 * a * b).div(den)
    D = D.masked_fill(~pos, 0)
    it = torch.tensor(n_iter, dtype=torch.long)
    due = nxt <= it + 1  # :251
    nxt[due] += per[due]  # :253-255
    D = D.masked_fill(~due, 0)
    g_att = torch.einsum("ijk,ij->ik", Zq.unsqueeze(1) - Zk, D).clamp(-4, 4)  # :258-263
    # --- repulsion, umap.py:266-292
    Zn = Z[neg.long()]
    R = _sq_to_rows(Zq, Zn)
    den = 1 + a * R**b
    R = R.add(eps).mul(den).reciprocal().mul(-2 * b)  # :274-276
    quota = (due.sum(dim=1) * negative_sample_rate).to(torch.long)  # :279-281
    col = torch.arange(neg.shape[1])
    R = R.masked_fill(col[None, :].ge(quota[:, None]), 0)  # :282-284
    g_rep = torch.einsum("ijk,ij->ik", Zq.unsqueeze(1) - Zn, R).clamp(-4, 4)  # :286-291
    return lam * g_att + repulsion * g_rep


def umap_run(Z0, idx, per, nxt0, negs, lrs, a, b, n_iter0=0, bounds=None,
             negative_sample_rate=5, return_all=False):
    """Loop of ``affinity_matcher.py:308-430`` (closed-form branch, plain SGD).

    ``negs[t]`` is the (already adjusted) negative table of step t — one
    ``[N, n_neg]`` tensor, or a list of per-rank chunk tables when ``bounds``
    (list of (start, end)) describes a row partition (distributed semantics,
    ``affinity_matcher.py:395-413``: all chunk gradients come from the old Z).
    """
    Z = Z0.clone()
    nxt = nxt0.clone()
    n = Z.shape[0]
    bounds = [(0, n)] if bounds is None else bounds
    traj = []
    for t, lr in enumerate(lrs):
        G = torch.zeros_like(Z)
        for r, (s, e) in enumerate(bounds):
            neg_t = negs[t] if len(bounds) == 1 and not isinstance(negs[t], (list, tuple)) else negs[t][r]
            G[s:e] = umap_step(Z, idx[s:e], per[s:e], nxt[s:e], neg_t, n_iter0 + t, a, b,
                               chunk_start=s, chunk_size=e - s,
                               negative_sample_rate=negative_sample_rate)
        Z.add_(G, alpha=-float(lr))  # torch.optim.SGD foreach step, no momentum (umap.py:139)
        if return_all:
            traj.append(Z.clone())
    return (Z, nxt, traj) if return_all else (Z, nxt)
