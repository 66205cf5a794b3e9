"""Test infrastructure: a CPU stand-in for ``torchdr_b200.ops`` built on the oracle.

It lets the HOST code of the package (estimators, affinities, the optimisation-loop bookkeeping) run in the CPU
suite, so that its orchestration can be held against the reference's golden runs without a GPU.  It is never
importable from the product (tests/test_layout.py) and implements only what UMAP needs.
"""

import numpy as np
import torch

import oracle


def install(monkeypatch):
    """Point the package at the CPU stand-ins (and let CPU tensors through the device checks)."""
    from torchdr_b200 import _lib, affinity, distance, neighbor_embedding, ops

    def to_cpu_tensor(X, device="auto"):
        return torch.from_numpy(X) if isinstance(X, np.ndarray) else X

    for mod in (distance, affinity, neighbor_embedding):
        monkeypatch.setattr(mod, "_to_device_tensor", to_cpu_tensor)
    monkeypatch.setattr(_lib, "require_device", lambda device: (0, 10, 0))
    monkeypatch.setattr(torch.cuda, "synchronize", lambda *a, **k: None)
    for name, fn in (("knn_umap_fused", knn_umap_fused), ("symmetrize_csr", symmetrize_csr), ("max_value", max_value),
                     ("umap_schedule", umap_schedule), ("umap_compact", umap_compact), ("umap_step", umap_step),
                     ("umap_run", umap_run), ("csr_to_ell", csr_to_ell)):
        monkeypatch.setattr(ops, name, fn)


def _relabel(C, I, labels):
    """ids -> labels, distance ties inside a row ranked by label (what the kernel's label keys do)."""
    if labels is None:
        return C, I
    L = labels.long()[I.long()]
    o1 = torch.argsort(L, dim=1, stable=True)
    o2 = torch.argsort(C.gather(1, o1), dim=1, stable=True)
    order = o1.gather(1, o2)
    return C.gather(1, order), L.gather(1, order).to(I.dtype)


def knn_umap_fused(Xq, Xdb, k, q_row0=0, exclude_self=True, max_iter=100, want_dist=True, path="auto", prune=None,
                   sweep_stats=None, labels=None):
    assert Xq.shape[0] == Xdb.shape[0] and q_row0 == 0 and exclude_self
    C, I = oracle.knn_dense(Xdb, k)
    C, I = _relabel(C, I, labels)
    P, rho, sigma = oracle.umap_affinity_rows(C, k, max_iter=max_iter)
    return C, I, P, rho, sigma


def symmetrize_csr(Pm, idx, row0, n_total, ext=None, transpose_local=True, mode="sum_minus_prod"):
    assert ext is None and row0 == 0
    V, J = oracle.symmetrize_ell(Pm, idx, mode=mode)
    return oracle.ell_to_csr(V, J)


def csr_to_ell(rowptr, col, val, pad_val=0.0):
    return oracle.csr_to_ell(rowptr, col, val, pad_val=pad_val)


def max_value(val):
    return val.max().reshape(1)


def umap_schedule(val, a_max, max_iter):
    per, nxt = oracle.umap_edge_schedule(val.unsqueeze(0), max_iter)  # elementwise: the layout does not matter
    return per.squeeze(0), nxt.squeeze(0)


def umap_compact(rowptr, col, eps):
    live = torch.isfinite(eps)
    n = rowptr.numel() - 1
    row = torch.repeat_interleave(torch.arange(n), rowptr[1:] - rowptr[:-1])
    cnt = torch.zeros(n, dtype=torch.long).index_add_(0, row[live], torch.ones(int(live.sum()), dtype=torch.long))
    rp = torch.zeros(n + 1, dtype=torch.long)
    rp[1:] = cnt.cumsum(0)
    return rp, col[live].contiguous(), eps[live].contiguous(), eps[live].clone()


def _ell(rowptr, col, eps, eons):
    V, J = oracle.csr_to_ell(rowptr, col, eps, pad_val=float("inf"))
    N, _ = oracle.csr_to_ell(rowptr, col, eons, pad_val=float("inf"))
    return J, V, N


def umap_step(Z_in, Z_out, row0, n_local, rowptr, col, eps, eons, n_iter, a, b, lr, neg=None, n_neg=75, rate=5,
              seed=0, lam=1.0, repulsion=1.0, precise=False, grad_out=None, gnorm_sq=None, nan_flag=None, stats=None):
    assert row0 == 0 and n_local == Z_in.shape[0]
    if neg is None:  # the kernel would draw in place; the stand-in draws its own table (distribution only)
        n = Z_in.shape[0]
        g = torch.Generator().manual_seed(int(seed) * 7919 + int(n_iter))
        neg = oracle.adjust_negatives(torch.randint(0, n - 1, (n, n_neg), generator=g), torch.arange(n))
    J, per, nxt = _ell(rowptr, col, eps, eons)
    G = oracle.umap_step(Z_in, J, per, nxt, neg, n_iter, a, b, negative_sample_rate=rate, lam=lam, repulsion=repulsion)
    # write the advanced edge state back into the CSR array (row-major ELL order == CSR order)
    eons.copy_(nxt[J >= 0])
    Z_out.copy_(Z_in.add(G, alpha=-float(lr)))
    if grad_out is not None:
        grad_out.copy_(G)
    if gnorm_sq is not None:
        gnorm_sq += float((G.double() ** 2).sum())


def umap_run(Z_a, Z_b, rowptr, col, eps, eons, n_iter0, lrs, a, b, n_neg=75, rate=5, seed=0, lam=1.0, repulsion=1.0,
             precise=False, gnorm_sq=None, nan_flag=None, stats=None, sync=None):
    g = torch.Generator().manual_seed(int(seed) * 7919 + int(n_iter0))
    src, dst = Z_a, Z_b
    n = Z_a.shape[0]
    for t, lr in enumerate(lrs):
        neg = oracle.adjust_negatives(torch.randint(0, n - 1, (n, n_neg), generator=g), torch.arange(n))
        umap_step(src, dst, 0, n, rowptr, col, eps, eons, n_iter0 + t, a, b, lr, neg=neg, n_neg=n_neg, rate=rate,
                  lam=lam, repulsion=repulsion, gnorm_sq=gnorm_sq if t == len(lrs) - 1 else None)
        src, dst = dst, src
    return src


# ---- entropic-affinity estimators (LargeVis, TSNE): gradients by autograd of the oracle's losses, as in the reference
def knn(Xq, Xdb, k, q_row0=0, exclude_self=True, metric="sqeuclidean", path="auto", prune=None, sweep_stats=None,
        labels=None):
    if Xq is Xdb or (Xq.shape == Xdb.shape and Xq.data_ptr() == Xdb.data_ptr()):
        return _relabel(*oracle.knn_dense(Xdb, k, metric, exclude_self), labels)
    # cross / chunk queries: distance/torch.py:81-122 on (Xq, Xdb), self excluded by global id
    C = oracle.pairwise_full(Xq, Xdb, metric)
    if exclude_self:
        r = torch.arange(Xq.shape[0])
        C[r, r + q_row0] += 1e12
    v, i = C.topk(k, dim=1, largest=False)
    return v, i.int()


def pairwise_full(X, Y=None, metric="sqeuclidean", exclude_diag=False, path="auto"):
    return oracle.pairwise_full(X, Y, metric, exclude_diag)


def indexed_distances(X, key_idx, Y=None, query_idx=None, metric="sqeuclidean"):
    Y = X if Y is None else Y
    Xq = X if query_idx is None else X[query_idx.long()]
    D = torch.sum((Xq.unsqueeze(1) - Y[key_idx.long()]) ** 2, dim=-1)  # distance/base.py:384-385
    return D.sqrt() if metric == "euclidean" else D


def entropic_affinity_rows(C, target_entropy, log_n_total, bounds=None, max_iter=100):
    perp = int(round(float(np.exp(target_entropy - 1))))
    n_total = int(round(float(np.exp(log_n_total))))
    return oracle.entropic_affinity_rows(C, perp, n_total=n_total, max_iter=max_iter, use_bounds=bounds is not None)


def _own_negatives(n, n_neg, seed, n_iter):
    g = torch.Generator().manual_seed(int(seed) * 7919 + int(n_iter))
    return oracle.adjust_negatives(torch.randint(0, n - 1, (n, n_neg), generator=g), torch.arange(n))


def largevis_grad(Z, row0, n_local, Pm, idx, grad, n_iter, neg=None, n_neg=5, seed=0, lam=1.0, repulsion=1.0):
    assert row0 == 0 and n_local == Z.shape[0]
    neg = _own_negatives(n_local, n_neg, seed, n_iter) if neg is None else neg
    from oracle.largevis import largevis_loss

    Zp = Z.detach().clone().requires_grad_(True)
    largevis_loss(Zp, Pm, idx, neg, torch.arange(n_local), Z.shape[0], lam, repulsion).backward()
    grad += Zp.grad


def largevis_step(Z_in, Z_out, row0, n_local, rowptr, col, val, grad_scratch, mom, n_iter, lr, momentum, first, n_neg=5,
                  seed=0, lam=1.0, repulsion=1.0, gnorm_sq=None, nan_flag=None, peer_ptrs=()):
    """tdr_largevis_step_f32 stand-in: gradient of the local rows from the UNION graph S = P + P^T (CSR) — attraction
    2 lam S_ij Q_ij (z_i - z_j) — plus both halves of every negative pair of the stand-in's own stream (pull on the
    sampling row, push on the sampled row), then torch.optim.SGD-with-momentum arithmetic on the local rows."""
    n = Z_in.shape[0]
    Z = Z_in.double()
    G = torch.zeros(n, 2, dtype=torch.float64)
    cnt = (rowptr[1:] - rowptr[:-1]).long()
    rows = torch.repeat_interleave(torch.arange(n_local) + row0, cnt)
    cj = col.long()
    d = Z[rows] - Z[cj]
    Q = 1.0 / (2.0 + (d * d).sum(1))
    G.index_add_(0, rows, (2.0 * lam * val.double() * Q).unsqueeze(1) * d)
    neg = _own_negatives(n, n_neg, seed, n_iter)  # all rows' negatives: pull for local i, push onto local j
    i_all = torch.arange(n).repeat_interleave(n_neg)
    j_all = neg.reshape(-1).long()
    dn = Z[i_all] - Z[j_all]
    Qn = 1.0 / (2.0 + (dn * dn).sum(1))
    c = (-2.0 * repulsion / n * Qn * Qn / (1.0 - Qn)).unsqueeze(1) * dn
    G.index_add_(0, i_all, c)
    G.index_add_(0, j_all, -c)
    g = G[row0:row0 + n_local].float()
    buf = g.clone() if first else mom * momentum + g
    mom.copy_(buf)
    Z_out[row0:row0 + n_local] = Z_in[row0:row0 + n_local] - lr * buf
    if gnorm_sq is not None:
        gnorm_sq += float((g.double() ** 2).sum())


def tsne_workspace(n_local, device):
    return torch.zeros(8, dtype=torch.uint8)


def tsne_grad(Z, row0, n_local, Pm, idx, lam, phase, grad, ws, repulsion=1.0):
    if phase == 0:
        return
    from oracle.tsne import tsne_loss

    Zp = Z.detach().clone().requires_grad_(True)
    tsne_loss(Zp, Pm, idx, torch.arange(n_local), lam, repulsion).backward()
    grad += Zp.grad


def sgd_momentum(Z, buf, grad, lr, momentum, first, gnorm_sq=None, nan_flag=None):
    # torch.optim.SGD: buf = grad (first use) | momentum * buf + grad ; param -= lr * buf
    if first:
        buf.copy_(grad)
    else:
        buf.mul_(momentum).add_(grad)
    Z.add_(buf, alpha=-lr)
    if gnorm_sq is not None:
        gnorm_sq += float((grad.double() ** 2).sum())


def infotsne_grad(Z, row0, n_local, Pm, idx, grad, n_iter, neg=None, n_neg=300, seed=0, lam=1.0, repulsion=1.0):
    assert row0 == 0 and n_local == Z.shape[0]
    neg = _own_negatives(n_local, n_neg, seed, n_iter) if neg is None else neg
    Zp = Z.detach().clone().requires_grad_(True)
    oracle.infotsne_loss(Zp, Pm, idx, neg, torch.arange(n_local), Z.shape[0], lam, repulsion).backward()
    grad += Zp.grad


def sne_grad(Z, row0, n_local, Pm, idx, lam, repulsion, phase, grad, row_sums):
    if phase == 0:
        return
    Zp = Z.detach().clone().requires_grad_(True)
    oracle.sne_loss(Zp, Pm, idx, torch.arange(n_local), lam, repulsion).backward()
    grad += Zp.grad


def install_entropic(monkeypatch):
    from torchdr_b200 import ops

    for name, fn in (("knn", knn), ("pairwise_full", pairwise_full), ("indexed_distances", indexed_distances),
                     ("entropic_affinity_rows", entropic_affinity_rows), ("largevis_grad", largevis_grad),
                     ("largevis_step", largevis_step), ("tsne_workspace", tsne_workspace), ("tsne_grad", tsne_grad), ("infotsne_grad", infotsne_grad),
                     ("sne_grad", sne_grad), ("sgd_momentum", sgd_momentum)):
        monkeypatch.setattr(ops, name, fn)


# ---- row-sharded stand-ins (world_size > 1 under gloo): same semantics as the CUDA entry points for a row chunk
def install_sharded(monkeypatch):
    from torchdr_b200 import distributed, neighbor_embedding, ops

    monkeypatch.setattr(torch.cuda, "set_device", lambda *a, **k: None)
    # the real sharded upload (own rows copied, the rest all-gathered), on CPU tensors over gloo
    monkeypatch.setattr(neighbor_embedding, "upload_sharded", lambda X, device: distributed.upload_sharded(X, "cpu"))
    for name, fn in (("knn_umap_fused", knn_umap_fused_chunk), ("symmetrize_export", symmetrize_export),
                     ("symmetrize_csr", symmetrize_csr_chunk), ("umap_step", umap_step_chunk)):
        monkeypatch.setattr(ops, name, fn)


def knn_umap_fused_chunk(Xq, Xdb, k, q_row0=0, exclude_self=True, max_iter=100, want_dist=True, path="auto", prune=None,
                         sweep_stats=None, labels=None):
    C, I = oracle.knn_dense(Xdb, k)  # full problem, then this rank's rows: identical values for every partition
    P, rho, sigma = oracle.umap_affinity_rows(C, k, max_iter=max_iter)
    s, e = q_row0, q_row0 + Xq.shape[0]
    return C[s:e].contiguous(), I[s:e].contiguous(), P[s:e].contiguous(), rho[s:e].contiguous(), sigma[s:e].contiguous()


def symmetrize_export(Pm, idx, row0, n_total, world, rank):
    """tdr_symmetrize_export_f32: edge (i -> j, v) of a local row i whose j lives on another rank is sent to owner(j)
    as the transposed entry (row = j, col = i, val = v), packed by destination rank."""
    n_local, k = Pm.shape
    i = (torch.arange(n_local) + row0).repeat_interleave(k)
    j = idx.reshape(-1).long()
    v = Pm.reshape(-1)
    owner = torch.tensor([oracle.owner_of(int(x), n_total, world) for x in j.tolist()], dtype=torch.long)
    away = owner != rank
    order = torch.argsort(owner[away], stable=True)
    counts = torch.bincount(owner[away], minlength=world)
    return counts, j[away][order].contiguous(), i[away][order].int().contiguous(), v[away][order].contiguous()


def symmetrize_csr_chunk(Pm, idx, row0, n_total, ext=None, transpose_local=True, mode="sum_minus_prod"):
    """Q = P + P^T - P o P^T for the local rows; P^T entries come from local edges landing on local rows and from the
    received triples.  Same three separately rounded fp32 operations as utils/sparse.py:163-164."""
    n_local, k = Pm.shape
    A = torch.zeros(n_local, n_total)
    B = torch.zeros(n_local, n_total)
    hasA = torch.zeros(n_local, n_total, dtype=torch.bool)
    hasB = torch.zeros(n_local, n_total, dtype=torch.bool)
    r = torch.arange(n_local).repeat_interleave(k)
    j = idx.reshape(-1).long()
    A[r, j] = Pm.reshape(-1)
    hasA[r, j] = True
    local = (j >= row0) & (j < row0 + n_local)
    B[j[local] - row0, r[local] + row0] = Pm.reshape(-1)[local]
    hasB[j[local] - row0, r[local] + row0] = True
    if ext is not None and ext[0].numel() > 0:
        er, ec, ev = ext
        B[er - row0, ec.long()] = ev
        hasB[er - row0, ec.long()] = True
    Q = A + B - A * B
    pattern = hasA | hasB
    rowptr = torch.zeros(n_local + 1, dtype=torch.long)
    rowptr[1:] = pattern.sum(1).cumsum(0)
    rr, cc = pattern.nonzero(as_tuple=True)
    return rowptr, cc.int().contiguous(), Q[rr, cc].contiguous()


def umap_step_chunk(Z_in, Z_out, row0, n_local, rowptr, col, eps, eons, n_iter, a, b, lr, neg=None, n_neg=75, rate=5,
                    seed=0, lam=1.0, repulsion=1.0, precise=False, grad_out=None, gnorm_sq=None, nan_flag=None,
                    stats=None):
    assert neg is not None
    J, per, nxt = _ell(rowptr, col, eps, eons)
    G = oracle.umap_step(Z_in, J, per, nxt, neg, n_iter, a, b, chunk_start=row0, chunk_size=n_local,
                         negative_sample_rate=rate, lam=lam, repulsion=repulsion)
    eons.copy_(nxt[J >= 0])
    Z_out[row0:row0 + n_local] = Z_in[row0:row0 + n_local].add(G, alpha=-float(lr))
    if grad_out is not None:
        grad_out.copy_(G)
    if gnorm_sq is not None:
        gnorm_sq += float((G.double() ** 2).sum())
