"""Tensor-level wrappers over the C ABI (one function per entry point of include/tdrb200.h).

These take and return CUDA tensors; shapes/dtypes are validated here so that the C layer
only sees raw pointers.  No function here computes on the CPU.
"""

import ctypes

import numpy as np
import torch

from . import _lib
from ._lib import check, ptr, stream, workspace


def _dev_f32(x, name):
    if not isinstance(x, torch.Tensor):
        raise TypeError(f"{name} must be a torch.Tensor")
    _lib.require_device(x.device)
    if x.dtype != torch.float32:
        raise TypeError(f"[TorchDR-B200] {name} must be float32 (the path computes in fp32), got {x.dtype}")
    return x.contiguous()


def _prune_id(prune):
    if isinstance(prune, str):
        return _lib.KNN_PRUNE[prune]
    return -1 if prune is None else int(prune)


def _check_stats(stats):
    if stats is not None:
        assert stats.is_cuda and stats.dtype == torch.int64 and stats.numel() >= 2 and stats.is_contiguous()
    return stats


def _check_labels(labels, ndb):
    if labels is not None:
        assert labels.is_cuda and labels.dtype == torch.int32 and labels.numel() == ndb and labels.is_contiguous()
    return labels


def knn(Xq, Xdb, k, q_row0=0, exclude_self=True, metric="sqeuclidean", path="auto", prune=None, sweep_stats=None,
        labels=None):
    """Exact kNN of query rows against the database -> (dist[nq,k] f32, idx[nq,k] i32).

    ``path``: "auto" | "simt" | "tc"; ``prune``: None (default = on) | 0/"off" | 1/"on" | 2/"certified";
    ``sweep_stats``: optional int64[2] CUDA tensor accumulating (tiles swept, tiles of a full sweep);
    ``labels``: optional int32[ndb] — ids to report (and to rank distance ties by) instead of the row indices."""
    Xq, Xdb = _dev_f32(Xq, "Xq"), _dev_f32(Xdb, "Xdb")
    lib = _lib.load()
    nq, d = Xq.shape
    ndb = Xdb.shape[0]
    ws = workspace(lib.tdr_knn_workspace_bytes(nq, ndb, d, k), Xq.device)
    dist = torch.empty((nq, k), dtype=torch.float32, device=Xq.device)
    idx = torch.empty((nq, k), dtype=torch.int32, device=Xq.device)
    with torch.cuda.device(Xq.device):
        check(lib.tdr_knn_f32(ptr(Xq), nq, q_row0, ptr(Xdb), ndb, d, k, int(exclude_self), _lib.METRIC_IDS[metric],
                              ptr(dist), ptr(idx), _lib.KNN_PATHS[path], _prune_id(prune), ptr(_check_stats(sweep_stats)),
                              ptr(_check_labels(labels, ndb)), ptr(ws), ws.numel(), stream()), "tdr_knn_f32")
    return dist, idx


def knn_umap_fused(Xq, Xdb, k, q_row0=0, exclude_self=True, max_iter=100, want_dist=True, path="auto", prune=None,
                   sweep_stats=None, labels=None):
    """Fused kNN + UMAP rho/sigma search -> (dist|None, idx, P, rho, sigma).  Options as in ``knn``."""
    Xq, Xdb = _dev_f32(Xq, "Xq"), _dev_f32(Xdb, "Xdb")
    lib = _lib.load()
    nq, d = Xq.shape
    ndb = Xdb.shape[0]
    dev = Xq.device
    ws = workspace(lib.tdr_knn_workspace_bytes(nq, ndb, d, k), dev)
    dist = torch.empty((nq, k), dtype=torch.float32, device=dev) if want_dist else None
    idx = torch.empty((nq, k), dtype=torch.int32, device=dev)
    Pm = torch.empty((nq, k), dtype=torch.float32, device=dev)
    rho = torch.empty((nq,), dtype=torch.float32, device=dev)
    sigma = torch.empty((nq,), dtype=torch.float32, device=dev)
    with torch.cuda.device(dev):
        check(lib.tdr_knn_umap_fused_f32(ptr(Xq), nq, q_row0, ptr(Xdb), ndb, d, k, int(exclude_self), max_iter,
                                         ptr(dist), ptr(idx), ptr(Pm), ptr(rho), ptr(sigma), _lib.KNN_PATHS[path],
                                         _prune_id(prune), ptr(_check_stats(sweep_stats)), ptr(_check_labels(labels, ndb)),
                                         ptr(ws), ws.numel(), stream()), "tdr_knn_umap_fused_f32")
    return dist, idx, Pm, rho, sigma


def pairwise_full(X, Y=None, metric="sqeuclidean", exclude_diag=False, path="auto"):
    X = _dev_f32(X, "X")
    Y = X if Y is None or Y is X else _dev_f32(Y, "Y")
    lib = _lib.load()
    n, d = X.shape
    m = Y.shape[0]
    ws = workspace(lib.tdr_knn_workspace_bytes(n, m, d, 1), X.device)
    C = torch.empty((n, m), dtype=torch.float32, device=X.device)
    with torch.cuda.device(X.device):
        check(lib.tdr_pairwise_full_f32(ptr(X), n, ptr(Y), m, d, _lib.METRIC_IDS[metric], int(exclude_diag), ptr(C),
                                        _lib.KNN_PATHS[path], ptr(ws), ws.numel(), stream()), "tdr_pairwise_full_f32")
    return C


def tree_assign(X, rows, node_of_row, centres, cnorm):
    """Nearest centre of every listed row within its own tree node (reorder.py) -> int64 [m]."""
    X = _dev_f32(X, "X")
    rows, node_of_row = rows.contiguous(), node_of_row.contiguous()
    centres, cnorm = centres.contiguous(), cnorm.contiguous()
    assert rows.dtype == torch.int64 and node_of_row.dtype == torch.int64 and centres.dtype == torch.float32
    m = rows.numel()
    out = torch.empty(m, dtype=torch.int64, device=X.device)
    with torch.cuda.device(X.device):
        check(_lib.load().tdr_tree_assign_f32(ptr(X), X.shape[1], ptr(rows), ptr(node_of_row), m, ptr(centres), ptr(cnorm),
                                              centres.shape[1], ptr(out), stream()), "tdr_tree_assign_f32")
    return out


def tree_accumulate(X, rows, node_of_row, child, n_nodes, branch):
    """Sums [n_nodes * branch, d] and counts [n_nodes * branch] of the rows of every (node, child) (reorder.py)."""
    X = _dev_f32(X, "X")
    rows, node_of_row, child = rows.contiguous(), node_of_row.contiguous(), child.contiguous()
    assert rows.dtype == node_of_row.dtype == child.dtype == torch.int64
    d = X.shape[1]
    spare = _lib.TDR_TREE_SPARE_NODES
    sums = torch.empty(((n_nodes + spare) * branch, d), dtype=torch.float32, device=X.device)
    cnt = torch.empty(((n_nodes + spare) * branch,), dtype=torch.float32, device=X.device)
    with torch.cuda.device(X.device):
        check(_lib.load().tdr_tree_accumulate_f32(ptr(X), d, ptr(rows), ptr(node_of_row), ptr(child), rows.numel(), n_nodes,
                                                  branch, ptr(sums), ptr(cnt), stream()), "tdr_tree_accumulate_f32")
    return sums[:n_nodes * branch], cnt[:n_nodes * branch]


def umap_affinity_rows(C, max_iter=100):
    C = _dev_f32(C, "C")
    n, k = C.shape
    Pm = torch.empty_like(C)
    rho = torch.empty((n,), dtype=torch.float32, device=C.device)
    sigma = torch.empty((n,), dtype=torch.float32, device=C.device)
    with torch.cuda.device(C.device):
        check(_lib.load().tdr_umap_affinity_f32(ptr(C), n, k, max_iter, ptr(Pm), ptr(rho), ptr(sigma), stream()),
              "tdr_umap_affinity_f32")
    return Pm, rho, sigma


def entropic_affinity_rows(C, target_entropy, log_n_total, bounds=None, max_iter=100):
    """bounds = (b_num, b_den, b_lr, b_logp1) fp32 scalars or None (multi-GPU rule)."""
    C = _dev_f32(C, "C")
    n, k = C.shape
    logP = torch.empty_like(C)
    eps = torch.empty((n,), dtype=torch.float32, device=C.device)
    log_norm = torch.empty((n,), dtype=torch.float32, device=C.device)
    b = bounds if bounds is not None else (0.0, 0.0, 0.0, 0.0)
    with torch.cuda.device(C.device):
        check(_lib.load().tdr_entropic_affinity_f32(ptr(C), n, k, float(target_entropy), float(log_n_total),
                                                    int(bounds is not None), float(b[0]), float(b[1]), float(b[2]),
                                                    float(b[3]), max_iter, ptr(logP), ptr(eps), ptr(log_norm),
                                                    stream()), "tdr_entropic_affinity_f32")
    return logP, eps, log_norm


def indexed_distances(X, key_idx, Y=None, query_idx=None, metric="sqeuclidean"):
    """out[i, s] = dist(X[query_idx[i]], Y[key_idx[i, s]]) (exact-difference form)."""
    X = _dev_f32(X, "X")
    Y = X if Y is None else _dev_f32(Y, "Y")
    key_idx = key_idx.contiguous()
    assert key_idx.dtype in (torch.int32, torch.int64) and key_idx.dim() == 2
    nq, k = key_idx.shape
    if query_idx is not None:
        query_idx = query_idx.to(torch.int64).contiguous()
        assert query_idx.numel() == nq
    out = torch.empty((nq, k), dtype=torch.float32, device=X.device)
    with torch.cuda.device(X.device):
        check(_lib.load().tdr_indexed_dist_f32(ptr(X), ptr(query_idx), nq, ptr(Y), Y.shape[0], X.shape[1], ptr(key_idx),
                                               int(key_idx.dtype == torch.int64), k, _lib.METRIC_IDS[metric], ptr(out),
                                               stream()), "tdr_indexed_dist_f32")
    return out


def entropic_dense_rows(C, target_entropy, log_n_total, bounds=None, max_iter=100, inplace=True, want_logP=True):
    """Dense rows: C [n_rows, m] full distances -> (logP [n_rows, m] or None, eps, log_norm).

    ``inplace`` overwrites C with log_P (the matrix is 4 N^2 bytes: 40 GB at N = 100 k)."""
    C = _dev_f32(C, "C")
    n_rows, m = C.shape
    eps = torch.empty((n_rows,), dtype=torch.float32, device=C.device)
    log_norm = torch.empty((n_rows,), dtype=torch.float32, device=C.device)
    logP = (C if inplace else torch.empty_like(C)) if want_logP else None
    b = bounds if bounds is not None else (0.0, 0.0, 0.0, 0.0)
    with torch.cuda.device(C.device):
        check(_lib.load().tdr_entropic_dense_f32(ptr(C), n_rows, m, float(target_entropy), float(log_n_total),
                                                 int(bounds is not None), float(b[0]), float(b[1]), float(b[2]),
                                                 float(b[3]), max_iter, ptr(logP), ptr(eps), ptr(log_norm), stream()),
              "tdr_entropic_dense_f32")
    return logP, eps, log_norm


def symmetrize_csr(Pm, idx, row0, n_total, ext=None, transpose_local=True, mode="sum_minus_prod"):
    """Q = P + P^T - P o P^T (mode "sum_minus_prod") or P + P^T ("sum") for the local rows
    -> (rowptr i64[n+1], col i32[nnz], val f32[nnz])."""
    Pm = _dev_f32(Pm, "P")
    idx = idx.contiguous()
    assert idx.dtype == torch.int32 and idx.shape == Pm.shape
    lib = _lib.load()
    dev = Pm.device
    n_local, k = Pm.shape
    if ext is not None and ext[0].numel() > 0:
        er, ec, ev = (t.contiguous() for t in ext)
        assert er.dtype == torch.int64 and ec.dtype == torch.int32 and ev.dtype == torch.float32
        n_ext = er.numel()
    else:
        er = ec = ev = None
        n_ext = 0
    cap = 2 * n_local * k + n_ext
    ws = workspace(lib.tdr_symmetrize_workspace_bytes(n_local, k, n_ext), dev)
    rowptr = torch.empty((n_local + 1,), dtype=torch.int64, device=dev)
    col = torch.empty((cap,), dtype=torch.int32, device=dev)
    val = torch.empty((cap,), dtype=torch.float32, device=dev)
    nnz = torch.zeros((1,), dtype=torch.int64, device=dev)
    with torch.cuda.device(dev):
        check(lib.tdr_symmetrize_csr_f32(ptr(Pm), ptr(idx), n_local, k, row0, n_total, ptr(er), ptr(ec), ptr(ev),
                                         n_ext, int(transpose_local), _lib.SYM_MODES[mode], ptr(rowptr), ptr(col), ptr(val), ptr(nnz),
                                         ptr(ws), ws.numel(), stream()), "tdr_symmetrize_csr_f32")
    m = int(nnz.item())
    return rowptr, col[:m].clone(), val[:m].clone()


def symmetrize_export(Pm, idx, row0, n_total, world, rank):
    """Transposed edges owned by other ranks, packed by destination -> (counts[world], row, col, val)."""
    Pm = _dev_f32(Pm, "P")
    idx = idx.contiguous()
    dev = Pm.device
    n_local, k = Pm.shape
    counts = torch.zeros((2 * world,), dtype=torch.int64, device=dev)
    row = torch.empty((n_local * k,), dtype=torch.int64, device=dev)
    col = torch.empty((n_local * k,), dtype=torch.int32, device=dev)
    val = torch.empty((n_local * k,), dtype=torch.float32, device=dev)
    with torch.cuda.device(dev):
        check(_lib.load().tdr_symmetrize_export_f32(ptr(Pm), ptr(idx), n_local, k, row0, n_total, world, rank,
                                                    ptr(counts), ptr(row), ptr(col), ptr(val), stream()),
              "tdr_symmetrize_export_f32")
    c = counts[:world].cpu()
    tot = int(c.sum())
    return c, row[:tot], col[:tot], val[:tot]


def csr_to_ell(rowptr, col, val, pad_val=0.0):
    n_local = rowptr.numel() - 1
    width = int((rowptr[1:] - rowptr[:-1]).max().item()) if n_local > 0 else 0
    ev = torch.empty((n_local, width), dtype=torch.float32, device=val.device)
    ei = torch.empty((n_local, width), dtype=torch.int64, device=val.device)
    with torch.cuda.device(val.device):
        check(_lib.load().tdr_csr_to_ell_f32(ptr(rowptr), ptr(col), ptr(val), n_local, width, float(pad_val), ptr(ev),
                                             ptr(ei), stream()), "tdr_csr_to_ell_f32")
    return ev, ei


def max_value(val):
    out = torch.zeros((1,), dtype=torch.float32, device=val.device)
    with torch.cuda.device(val.device):
        check(_lib.load().tdr_max_f32(ptr(val), val.numel(), ptr(out), stream()), "tdr_max_f32")
    return out


def umap_schedule(val, a_max, max_iter):
    eps = torch.empty_like(val)
    eons = torch.empty_like(val)
    with torch.cuda.device(val.device):
        check(_lib.load().tdr_umap_schedule_f32(ptr(val), val.numel(), float(a_max), int(max_iter), ptr(eps),
                                                ptr(eons), stream()), "tdr_umap_schedule_f32")
    return eps, eons


def umap_compact(rowptr, col, eps):
    lib = _lib.load()
    dev = eps.device
    n_local = rowptr.numel() - 1
    nnz = eps.numel()
    ws = workspace(lib.tdr_compact_workspace_bytes(n_local, nnz), dev)
    orp = torch.empty_like(rowptr)
    ocol = torch.empty_like(col)
    oeps = torch.empty_like(eps)
    oeons = torch.empty_like(eps)
    cnt = torch.zeros((1,), dtype=torch.int64, device=dev)
    with torch.cuda.device(dev):
        check(lib.tdr_umap_compact_f32(ptr(rowptr), ptr(col), ptr(eps), n_local, nnz, ptr(orp), ptr(ocol), ptr(oeps),
                                       ptr(oeons), ptr(cnt), ptr(ws), ws.numel(), stream()), "tdr_umap_compact_f32")
    m = int(cnt.item())
    return orp, ocol[:m].clone(), oeps[:m].clone(), oeons[:m].clone()


def umap_step(Z_in, Z_out, row0, n_local, rowptr, col, eps, eons, n_iter, a, b, lr, neg=None, n_neg=75, rate=5,
              seed=0, lam=1.0, repulsion=1.0, precise=False, grad_out=None, gnorm_sq=None, nan_flag=None, stats=None):
    n_total = Z_in.shape[0]
    if neg is not None:
        assert neg.dtype == torch.int64 and neg.is_contiguous() and neg.shape == (n_local, n_neg)
    with torch.cuda.device(Z_in.device):
        check(_lib.load().tdr_umap_step_f32(ptr(Z_in), ptr(Z_out), n_total, row0, n_local, ptr(rowptr), ptr(col),
                                            ptr(eps), ptr(eons), ptr(neg), n_neg, rate, seed, n_iter, float(a),
                                            float(b), float(lam), float(repulsion), float(lr), int(precise),
                                            ptr(grad_out), ptr(gnorm_sq), ptr(nan_flag), ptr(stats), stream()),
              "tdr_umap_step_f32")


def umap_step_p2p(Z_in, Z_out, row0, n_local, rowptr, col, eps, eons, n_iter, a, b, lr, peer_ptrs, n_neg=75, rate=5,
                  seed=0, lam=1.0, repulsion=1.0, gnorm_sq=None, nan_flag=None):
    """Step + exchange in one kernel: updated rows are also stored into the peers' Z_out (device addresses)."""
    arr = (ctypes.c_uint64 * max(len(peer_ptrs), 1))(*[int(x) for x in peer_ptrs])
    with torch.cuda.device(Z_in.device):
        check(_lib.load().tdr_umap_step_p2p_f32(ptr(Z_in), ptr(Z_out), Z_in.shape[0], row0, n_local, ptr(rowptr),
                                                ptr(col), ptr(eps), ptr(eons), n_neg, rate, seed, n_iter, float(a),
                                                float(b), float(lam), float(repulsion), float(lr), ptr(gnorm_sq),
                                                ptr(nan_flag), arr, len(peer_ptrs), stream()),
              "tdr_umap_step_p2p_f32")


class RunSync:
    """Device words of the persistent step kernel's barriers (``sync_words`` of tdr_umap_run_f32): zero-initialised
    once, reused by every call on the same embedding; ``epoch`` counts the iterations run on it."""

    def __init__(self, device):
        self.words = torch.zeros(_lib.TDR_RUN_SYNC_WORDS, dtype=torch.int32, device=device)
        self.epoch = 0

    def check(self):
        """Host-side check of the status word (synchronises): a peer or a CTA that never reached a barrier."""
        st = int(self.words[_lib.TDR_RUN_STATUS_WORD].item())
        if st != 0:
            who = "a peer GPU" if st == 1 else "a thread block of this GPU"
            raise _lib.B200EngineError(f"[TorchDR-B200] the persistent UMAP loop was aborted: {who} did not reach the "
                                       "iteration barrier within the time limit (is a rank dead?).")


def umap_run_p2p(peer, cur, row0, n_local, rowptr, col, eps, eons, n_iter0, lrs, a, b, n_neg=75, rate=5, seed=0,
                 lam=1.0, repulsion=1.0, gnorm_sq=None, nan_flag=None, stats=None, timeout_s=60.0):
    """len(lrs) sharded iterations in ONE persistent launch (step + NVLink peer stores + cross-GPU barrier in-kernel).

    ``peer`` is a distributed.PeerEmbedding, ``cur`` the index of the buffer holding the current embedding;
    returns the index of the buffer holding the result."""
    lrs = np.ascontiguousarray(lrs, dtype=np.float32)
    n = len(lrs)
    pa, pb = peer.ptr_arrays[cur], peer.ptr_arrays[1 - cur]
    Za, Zb = peer.bufs[cur], peer.bufs[1 - cur]
    with torch.cuda.device(Za.device):
        check(_lib.load().tdr_umap_run_p2p_f32(ptr(Za), ptr(Zb), Za.shape[0], row0, n_local, ptr(rowptr), ptr(col),
                                               ptr(eps), ptr(eons), n_neg, rate, seed, n_iter0, n,
                                               lrs.ctypes.data_as(ctypes.POINTER(ctypes.c_float)), float(a), float(b),
                                               float(lam), float(repulsion), ptr(gnorm_sq), ptr(nan_flag), ptr(stats),
                                               ptr(peer.sync.words), pa, pb, ptr(peer.flags), peer.flag_ptr_array,
                                               peer.world - 1, peer.rank, peer.world, peer.sync.epoch % (1 << 32),
                                               float(timeout_s), stream()), "tdr_umap_run_p2p_f32")
    peer.sync.epoch += n
    return cur if n % 2 == 0 else 1 - cur


def umap_run(Z_a, Z_b, rowptr, col, eps, eons, n_iter0, lrs, a, b, n_neg=75, rate=5, seed=0, lam=1.0,
             repulsion=1.0, precise=False, gnorm_sq=None, nan_flag=None, stats=None, sync=None):
    """len(lrs) iterations on one GPU (one persistent launch); returns the tensor holding the result.
    ``sync``: a RunSync kept by the caller across calls (a fresh one is made when omitted)."""
    lrs = np.ascontiguousarray(lrs, dtype=np.float32)
    n = len(lrs)
    if sync is None:
        sync = RunSync(Z_a.device)
    with torch.cuda.device(Z_a.device):
        check(_lib.load().tdr_umap_run_f32(ptr(Z_a), ptr(Z_b), Z_a.shape[0], ptr(rowptr), ptr(col), ptr(eps),
                                           ptr(eons), n_neg, rate, seed, n_iter0, n,
                                           lrs.ctypes.data_as(ctypes.POINTER(ctypes.c_float)), float(a), float(b),
                                           float(lam), float(repulsion), int(precise), ptr(gnorm_sq), ptr(nan_flag),
                                           ptr(stats), ptr(sync.words), sync.epoch % (1 << 32), stream()),
              "tdr_umap_run_f32")
    sync.epoch += n
    return Z_a if n % 2 == 0 else Z_b


def largevis_grad(Z, row0, n_local, Pm, idx, grad, n_iter, neg=None, n_neg=5, seed=0, lam=1.0, repulsion=1.0):
    with torch.cuda.device(Z.device):
        check(_lib.load().tdr_largevis_grad_f32(ptr(Z), Z.shape[0], row0, n_local, ptr(Pm), ptr(idx), Pm.shape[1],
                                                ptr(neg), n_neg, seed, n_iter, float(lam), float(repulsion),
                                                ptr(grad), stream()), "tdr_largevis_grad_f32")


def largevis_step(Z_in, Z_out, row0, n_local, rowptr, col, val, grad_scratch, mom, n_iter, lr, momentum, first, n_neg=5,
                  seed=0, lam=1.0, repulsion=1.0, gnorm_sq=None, nan_flag=None, peer_ptrs=()):
    """One row-local LargeVis iteration (gradient from the P + P^T graph, momentum SGD on the local rows, rows written to
    Z_out and to the NVLink peers' Z_out)."""
    arr = (ctypes.c_uint64 * max(len(peer_ptrs), 1))(*[int(x) for x in peer_ptrs])
    with torch.cuda.device(Z_in.device):
        check(_lib.load().tdr_largevis_step_f32(ptr(Z_in), ptr(Z_out), Z_in.shape[0], row0, n_local, ptr(rowptr), ptr(col),
                                                ptr(val), n_neg, seed, n_iter, float(lam), float(repulsion),
                                                ptr(grad_scratch), ptr(mom), float(lr), float(momentum), int(first),
                                                ptr(gnorm_sq), ptr(nan_flag), arr, len(peer_ptrs), stream()),
              "tdr_largevis_step_f32")


def tsne_workspace(n_local, device):
    return workspace(_lib.load().tdr_tsne_workspace_bytes(n_local), device)


def tsne_grad(Z, row0, n_local, Pm, idx, lam, phase, grad, ws, repulsion=1.0):
    k = Pm.shape[1] if Pm is not None else 0
    with torch.cuda.device(Z.device):
        check(_lib.load().tdr_tsne_grad_f32(ptr(Z), Z.shape[0], row0, n_local, ptr(Pm), ptr(idx), k, float(lam),
                                            float(repulsion), int(phase), ptr(grad), ptr(ws), ws.numel(), stream()),
              "tdr_tsne_grad_f32")


def infotsne_grad(Z, row0, n_local, Pm, idx, grad, n_iter, neg=None, n_neg=300, seed=0, lam=1.0, repulsion=1.0):
    with torch.cuda.device(Z.device):
        check(_lib.load().tdr_infotsne_grad_f32(ptr(Z), Z.shape[0], row0, n_local, ptr(Pm), ptr(idx), Pm.shape[1],
                                                ptr(neg), n_neg, seed, n_iter, float(lam), float(repulsion),
                                                ptr(grad), stream()), "tdr_infotsne_grad_f32")


def sne_grad(Z, row0, n_local, Pm, idx, lam, repulsion, phase, grad, row_sums):
    k = Pm.shape[1] if Pm is not None else 0
    with torch.cuda.device(Z.device):
        check(_lib.load().tdr_sne_grad_f32(ptr(Z), Z.shape[0], row0, n_local, ptr(Pm), ptr(idx), k, float(lam),
                                           float(repulsion), int(phase), ptr(grad), ptr(row_sums), stream()),
              "tdr_sne_grad_f32")


def sgd_momentum(Z, buf, grad, lr, momentum, first, gnorm_sq=None, nan_flag=None):
    with torch.cuda.device(Z.device):
        check(_lib.load().tdr_sgd_momentum_f32(ptr(Z), ptr(buf), ptr(grad), Z.numel(), float(lr), float(momentum),
                                               int(first), ptr(gnorm_sq), ptr(nan_flag), stream()),
              "tdr_sgd_momentum_f32")
