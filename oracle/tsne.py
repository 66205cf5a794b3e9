"""Oracle: t-SNE loss + momentum-SGD loop with early exaggeration (test infrastructure).

Restates ``torchdr/neighbor_embedding/tsne.py:162-180`` (attractive on the
directed kNN edges; repulsive = logsumexp over ALL N x N pairs, diagonal
included, of -log(1 + C) with C in the expanded form of
``distance/torch.py:89-91``), ``neighbor_embedding/base.py:282-343``
(early-exaggeration switch that rebuilds the optimiser; lr "auto"; momentum
0.5 while lambda > 1 else 0.8 — but see the param-group note in tsne_run) and ``affinity_matcher.py:414-429``.
"""

import torch


def tsne_loss(Z, P, idx, rows, lam, repulsion=1.0):
    Zq = Z[rows]
    D = torch.sum((Zq.unsqueeze(1) - Z[idx.long()]) ** 2, dim=-1)
    att = -(P * (-(1 + D).log())).sum()  # tsne.py:162-170
    nz = (Z**2).sum(-1)
    C = nz.unsqueeze(-1) + nz.unsqueeze(-2) - 2 * (Z @ Z.transpose(-1, -2))
    rep = (-(1 + C).log()).logsumexp((0, 1))  # tsne.py:172-180
    return lam * att + repulsion * rep  # NE base.py:223-242: early_exag * attractive + repulsion_strength * repulsive


def tsne_run(Z0, P, idx, n_steps, exag=12.0, exag_iter=250, lr=None, return_grads=False):
    n = Z0.shape[0]
    Z = torch.nn.Parameter(Z0.clone())
    rows = torch.arange(n)

    # affinity_matcher.py:588-590: params_ is ONE dict reused for every optimiser build.
    # torch's add_param_group fills it with setdefault(), so when the optimiser is rebuilt at
    # the end of early exaggeration the dict still carries the FIRST build's lr and momentum:
    # only the momentum buffer is reset (verified on the reference: lr stays 50, momentum 0.5
    # after the switch although lr_ becomes 75).  Restated as is.
    group = {"params": Z}

    def make_opt(lam):
        lr_ = max(n / lam / 4, 50) if lr is None else lr  # NE base.py:299-310
        mom = 0.5 if lam > 1 else 0.8  # NE base.py:331-338
        return torch.optim.SGD([group], lr=lr_, momentum=mom)

    lam = exag
    opt = make_opt(lam)
    grads = []
    for t in range(n_steps):
        opt.zero_grad(set_to_none=True)
        tsne_loss(Z, P, idx, rows, lam).backward()
        if return_grads:
            grads.append(Z.grad.detach().clone())
        opt.step()
        if lam > 1 and t == exag_iter:  # NE base.py:282-295
            lam = 1
            opt = make_opt(lam)
    out = Z.detach().clone()
    return (out, grads) if return_grads else out
