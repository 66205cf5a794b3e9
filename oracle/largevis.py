"""Oracle: LargeVis loss + momentum-SGD loop (test infrastructure).

Restates ``torchdr/neighbor_embedding/largevis.py:181-201`` (losses),
``neighbor_embedding/base.py:223-233`` (lambda * attractive + repulsive),
``:299-350`` (lr "auto" = max(N/lambda/4, 50), SGD momentum 0.8, LinearLR with
torch defaults because ``scheduler_kwargs=None``, largevis.py:118) and the
autograd branch of ``affinity_matcher.py:414-429``.  Gradients come from torch
autograd exactly as in the reference.
"""

import torch


def largevis_loss(Z, P, idx, neg, rows, n_total, lam=1.0, repulsion=1.0):
    # each loss gathers its own query rows (distance/base.py:336-339 is called twice);
    # sharing one gather would change the autograd accumulation order
    # largevis.py:192-201 ; distance/base.py:384-385
    D = torch.sum((Z[rows].unsqueeze(1) - Z[idx.long()]) ** 2, dim=-1)
    Q = 1.0 / (1.0 + D)
    Q = Q / (Q + 1)
    att = -(P * Q.log()).sum()  # utils/utils.py:121-124
    # largevis.py:181-190
    Dn = torch.sum((Z[rows].unsqueeze(1) - Z[neg.long()]) ** 2, dim=-1)
    Qn = 1.0 / (1.0 + Dn)
    Qn = Qn / (Qn + 1)
    rep = -((1 - Qn).log()).sum() / n_total
    return lam * att + repulsion * rep


def largevis_run(Z0, P, idx, negs, n_steps, lr=None, momentum=0.8, bounds=None,
                 return_grads=False):
    """``affinity_matcher.py:308-430`` autograd branch; ``negs[t]`` as in umap_run."""
    n = Z0.shape[0]
    Z = torch.nn.Parameter(Z0.clone())
    lr0 = max(n / 1.0 / 4, 50) if lr is None else lr  # NE base.py:308
    opt = torch.optim.SGD([Z], lr=lr0, momentum=momentum)  # NE base.py:331-343
    sch = torch.optim.lr_scheduler.LinearLR(opt)  # largevis.py:118 -> torch defaults
    bounds = [(0, n)] if bounds is None else bounds
    lrs, grads = [], []
    for t in range(n_steps):
        opt.zero_grad(set_to_none=True)
        lrs.append(float(opt.param_groups[0]["lr"]))
        total = None
        for r, (s, e) in enumerate(bounds):
            neg_t = negs[t] if len(bounds) == 1 and not isinstance(negs[t], (list, tuple)) else negs[t][r]
            part = largevis_loss(Z, P[s:e], idx[s:e], neg_t, torch.arange(s, e), n)
            total = part if total is None else total + part
        total.backward()
        if return_grads:
            grads.append(Z.grad.detach().clone())
        opt.step()
        sch.step()
    out = Z.detach().clone()
    return (out, lrs, grads) if return_grads else (out, lrs)
