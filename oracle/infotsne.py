"""Oracle: InfoTSNE loss + momentum-SGD loop with early exaggeration (test infrastructure).

Restates ``torchdr/neighbor_embedding/infotsne.py:179-197`` (attractive: cross entropy of P against
log Q = -log(1 + D) on the kNN rows; repulsive: per-row logsumexp of log Q over the sampled negatives,
summed and divided by N), ``neighbor_embedding/base.py:282-350`` (early-exaggeration switch rebuilding
optimiser AND scheduler; lr "auto"; momentum 0.5 / 0.8) and the autograd branch of
``affinity_matcher.py:414-429``.  Default scheduler: ``LinearLR`` with torch's defaults
(``infotsne.py:113-116``, ``scheduler_kwargs=None``).
"""

import torch


def infotsne_loss(Z, P, idx, neg, rows, n_total, lam=1.0, repulsion=1.0):
    # each loss gathers its own query rows (distance/base.py:336-339 is called twice)
    D = torch.sum((Z[rows].unsqueeze(1) - Z[idx.long()]) ** 2, dim=-1)  # distance/base.py:384-385
    att = -(P * (-(1 + D).log())).sum()  # infotsne.py:179-187, utils/utils.py:121-122
    Dn = torch.sum((Z[rows].unsqueeze(1) - Z[neg.long()]) ** 2, dim=-1)
    rep = (-(1 + Dn).log()).logsumexp(1).sum() / n_total  # infotsne.py:189-197
    return lam * att + repulsion * rep


def infotsne_run(Z0, P, idx, negs, n_steps, exag=12.0, exag_iter=250, lr=None, return_grads=False):
    n = Z0.shape[0]
    Z = torch.nn.Parameter(Z0.clone())
    rows = torch.arange(n)
    group = {"params": Z}  # ONE dict reused by every rebuild (affinity_matcher.py:588-590), see oracle/tsne.py

    def make_opt(lam):
        lr_ = max(n / lam / 4, 50) if lr is None else lr  # NE base.py:299-310
        mom = 0.5 if lam > 1 else 0.8  # NE base.py:331-338
        opt = torch.optim.SGD([group], lr=lr_, momentum=mom)
        return opt, torch.optim.lr_scheduler.LinearLR(opt)

    lam = exag
    opt, sch = make_opt(lam)
    lrs, grads = [], []
    for t in range(n_steps):
        opt.zero_grad(set_to_none=True)
        lrs.append(float(opt.param_groups[0]["lr"]))
        infotsne_loss(Z, P, idx, negs[t], rows, n, lam).backward()
        if return_grads:
            grads.append(Z.grad.detach().clone())
        opt.step()
        sch.step()
        if lam > 1 and t == exag_iter:  # NE base.py:282-295
            lam = 1
            opt, sch = make_opt(lam)
    out = Z.detach().clone()
    return (out, lrs, grads) if return_grads else (out, lrs)
