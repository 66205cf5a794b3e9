"""Oracle: SNE loss + momentum-SGD loop (test infrastructure).

Restates ``torchdr/neighbor_embedding/sne.py:162-179`` (attractive: sum P_ij D_ij on the directed kNN
edges; repulsive: per-row logsumexp of -C over ALL columns, diagonal included, C in the expanded form
of ``distance/torch.py:89-91``, summed and divided by N), ``neighbor_embedding/base.py:299-343``
(lr "auto" = max(N/4, 50), SGD momentum 0.8 — no early exaggeration by default, ``sne.py:116-117``) and
``affinity_matcher.py:414-429``.  No scheduler (``sne.py:102``).
"""

import torch


def sne_loss(Z, P, idx, rows, lam=1.0, repulsion=1.0):
    D = torch.sum((Z[rows].unsqueeze(1) - Z[idx.long()]) ** 2, dim=-1)  # distance/base.py:384-385
    att = -(P * (-D)).sum()  # sne.py:162-170
    nz = (Z**2).sum(-1)
    C = nz.unsqueeze(-1) + nz.unsqueeze(-2) - 2 * (Z @ Z.transpose(-1, -2))  # distance/torch.py:89-91
    rep = (-C).logsumexp(1).sum() / Z.shape[0]  # sne.py:172-176
    return lam * att + repulsion * rep


def sne_run(Z0, P, idx, n_steps, lr=None, momentum=0.8, return_grads=False):
    n = Z0.shape[0]
    Z = torch.nn.Parameter(Z0.clone())
    rows = torch.arange(n)
    lr0 = max(n / 1.0 / 4, 50) if lr is None else lr
    opt = torch.optim.SGD([Z], lr=lr0, momentum=momentum)
    grads = []
    for _ in range(n_steps):
        opt.zero_grad(set_to_none=True)
        sne_loss(Z, P, idx, rows).backward()
        if return_grads:
            grads.append(Z.grad.detach().clone())
        opt.step()
    out = Z.detach().clone()
    return (out, grads) if return_grads else out
