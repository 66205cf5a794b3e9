"""Affinity seam: ``UMAPAffinity`` and ``EntropicAffinity`` with the reference's call contract.

Mirrors ``torchdr/affinity/base.py:407-561`` (``SparseAffinity.__call__`` /
``SparseLogAffinity.__call__``), ``affinity/knn_normalized.py:335-496`` and
``affinity/entropic.py:118-312``.  Same constructor kwargs, same return values
(``values[n_local, W]`` + ``indices[n_local, W]`` int64 with -1 padding for the symmetrised UMAP
graph; ``log_P[n_local, k]`` + int32 indices for the entropic affinity), same fitted attributes
(``rho_``, ``eps_``, ``log_normalization_``, ``chunk_start_/chunk_end_/chunk_size_``).
Engine-native extra: ``csr_`` = (rowptr, col, val) of the symmetrised graph, which the UMAP
estimator consumes directly instead of the padded layout.
"""

import logging
import math
import os
import time

import torch
import torch.nn as nn

from . import _lib, ops
from .distance import _check_metric, _to_device_tensor
from .distributed import DistributedContext, exchange_edges, is_distributed


def check_neighbor_param(value, n_samples):
    """``torchdr/utils/validation.py:223-244``: long cast, clamp to [2, n-2]."""
    if n_samples <= 1:
        raise ValueError(f"[TorchDR] ERROR : Input has less than one sample : n_samples = {n_samples}.")
    return int(min(max(int(value), 2), n_samples - 2))


def _scalar_bisect_f32(f, lo, hi, max_iter):
    """1-element restatement of utils/root_search.py:17-77 on CPU fp32 tensors (host scalar setup only)."""
    tol = torch.tensor(1e-6)
    for _ in range(max_iter):
        if not bool(f(lo) > 0):
            break
        hi = torch.minimum(hi, lo)
        lo = lo * 0.5
    for _ in range(max_iter):
        if not bool(f(hi) < 0):
            break
        lo = torch.maximum(lo, hi)
        hi = hi * 2.0
    f_lo = f(lo)
    mid = (lo + hi) * 0.5
    f_mid = f(mid)
    for _ in range(max_iter):
        if not bool(f_mid.abs() >= tol):
            break
        if bool(f_mid * f_lo > 0):
            lo, f_lo = mid, f_mid
        else:
            hi = mid
        mid = (lo + hi) * 0.5
        f_mid = f(mid)
    return mid


def entropic_bound_scalars(n_rows, perplexity):
    """Data-independent fp32 scalars of the Vladymyrov bracket (affinity/entropic.py:68-106)."""
    tN = torch.tensor(float(n_rows), dtype=torch.float32)
    perp = torch.tensor(int(perplexity))
    cap = torch.minimum(torch.sqrt(2.0 * tN), perp)

    def gap(x):
        return torch.log(cap) - 2.0 * (1.0 - x) * torch.log(tN / (2.0 * (1.0 - x)))

    p1 = _scalar_bisect_f32(gap, torch.tensor(0.75), torch.tensor(1.0 - 1e-6), 1000)
    lr = torch.log(tN / perp)
    return (float(tN * lr), float(tN - 1), float(lr), float(torch.log((tN - 1) * p1 / (1.0 - p1))))


class _SparseAffinityBase(nn.Module):
    """Common plumbing of ``affinity/base.py:30-486`` (``Affinity(nn.Module, ABC)``; fitted quantities are
    non-persistent state dropped by ``clear_memory``)."""

    def __init__(self, metric="sqeuclidean", zero_diag=True, device="auto", backend=None, verbose=False,
                 compile=False, sparsity=True, distributed="auto", _pre_processed=False, knn_order="auto"):
        nn.Module.__init__(self)
        _check_metric(metric)
        if knn_order not in ("auto", "input", "tree", "presorted"):
            raise ValueError("[TorchDR-B200] knn_order must be 'auto', 'input' or 'tree'.")
        # engine-native option (not in the reference): row order the exact kNN search runs in — "auto" re-orders
        # inputs without index locality (torchdr_b200/reorder.py) and maps the result back; "presorted" is set by the
        # estimators, which run the whole fit in the tree order themselves
        self.knn_order = knn_order
        self.metric = metric
        self.zero_diag = zero_diag
        self.device = device
        self.backend = backend
        self.verbose = verbose
        self.compile = compile
        self.sparsity = sparsity
        self._pre_processed = _pre_processed
        self.logger = logging.getLogger(f"torchdr_b200.{self.__class__.__name__}")
        if self.verbose:
            self.logger.setLevel(logging.INFO)
        if distributed == "auto":  # affinity/base.py:323-326
            self.distributed = is_distributed()
        else:
            self.distributed = bool(distributed)
        if self.distributed:
            if not is_distributed():
                raise RuntimeError(
                    "[TorchDR] distributed=True requires launching with torchrun. "
                    "Example: torchrun --nproc_per_node=4 your_script.py"
                )
            if not self.sparsity:
                raise ValueError("[TorchDR] Distributed mode requires sparsity=True.")
            self.dist_ctx = DistributedContext()
            self.rank = self.dist_ctx.rank
            self.world_size = self.dist_ctx.world_size
            self.is_multi_gpu = self.world_size > 1
        else:
            self.dist_ctx = None
            self.rank = 0
            self.world_size = 1
            self.is_multi_gpu = False

    def _prepare(self, X):
        X = _to_device_tensor(X, self.device)
        return X.float().contiguous() if X.dtype != torch.float32 or not X.is_contiguous() else X

    def _search_order(self, X):
        """(permutation | None, prune mode) for the kNN search of this call (single-GPU only: the estimators handle
        the row-sharded case by running the whole fit in the tree order)."""
        if self.knn_order == "presorted":
            return None, "certified"
        if self.is_multi_gpu or self.knn_order == "input":
            return None, None
        from .reorder import choose_order

        perm = choose_order(X, self.knn_order)
        return perm, ("certified" if perm is not None else None)

    def _chunk(self, n):
        """Row chunk of this rank; records chunk_start_/end_/size_ (affinity/base.py:477-484)."""
        if self.distributed and self.dist_ctx is not None:
            s, e = self.dist_ctx.compute_chunk_bounds(n)
            self.chunk_start_, self.chunk_end_, self.chunk_size_ = s, e, e - s
            return s, e
        return 0, n

    def clear_memory(self):
        for name in ("rho_", "eps_", "log_normalization_", "csr_", "knn_"):
            if hasattr(self, name):
                delattr(self, name)


class UMAPAffinity(_SparseAffinityBase):
    """``torchdr/affinity/knn_normalized.py:335-496``."""

    def __init__(self, n_neighbors=30, max_iter=1000, sparsity=True, metric="sqeuclidean", zero_diag=True,
                 device="auto", backend=None, verbose=False, compile=False, symmetrize=True, distributed="auto",
                 _pre_processed=False, knn_order="auto"):
        self.n_neighbors = n_neighbors
        self.max_iter = max_iter
        self.symmetrize = symmetrize
        super().__init__(metric=metric, zero_diag=zero_diag, device=device, backend=backend, verbose=verbose,
                         compile=compile, sparsity=sparsity, distributed=distributed, _pre_processed=_pre_processed,
                         knn_order=knn_order)

    def compute_csr(self, X):
        """Engine-native result: kNN + sigma/rho search (one fused kernel) + symmetrisation -> CSR."""
        if not self.sparsity:
            raise NotImplementedError("[TorchDR-B200] UMAPAffinity(sparsity=False) is outside the accelerated path.")
        X = self._prepare(X)
        n = X.shape[0]
        k = check_neighbor_param(self.n_neighbors, n)
        if k > _lib.TDR_MAX_K:
            raise NotImplementedError(f"[TorchDR-B200] n_neighbors={k} exceeds the engine limit {_lib.TDR_MAX_K}.")
        if self.verbose:
            self.logger.info(f"Sparsity mode enabled, computing {k} nearest neighbors...")
        s, e = self._chunk(n)
        timing = os.environ.get("TDR_TIMING") == "1"  # stage timers of scripts/e2e_breakdown.py
        if timing:
            torch.cuda.synchronize()
            t0 = time.perf_counter()
        perm, prune = self._search_order(X)
        if perm is not None:
            # no index locality in the input: search in the tree order with the certified sweep.  The rows' own ids travel
            # as labels: the kernel reports them and ranks distance ties by them, so the result is the input-order
            # search's bit for bit; only the rows have to be put back in place
            from .reorder import unpermute_rows

            Xp = X[perm].contiguous()
            lab = perm.to(torch.int32)
            if self.metric == "sqeuclidean":
                dist, idx, P, rho, sigma = ops.knn_umap_fused(Xp, Xp, k, q_row0=0, exclude_self=bool(self.zero_diag),
                                                              max_iter=self.max_iter, prune=prune, labels=lab)
            else:
                dist, idx = ops.knn(Xp, Xp, k, q_row0=0, exclude_self=bool(self.zero_diag), metric=self.metric, prune=prune,
                                    labels=lab)
                P, rho, sigma = ops.umap_affinity_rows(dist, self.max_iter)
            idx, dist, P, rho, sigma = unpermute_rows(perm, idx, dist, P, rho, sigma)
            del Xp
        elif self.metric == "sqeuclidean":
            dist, idx, P, rho, sigma = ops.knn_umap_fused(X[s:e], X, k, q_row0=s, exclude_self=bool(self.zero_diag),
                                                          max_iter=self.max_iter, prune=prune)
        else:  # euclidean rows go through the two-kernel route
            dist, idx = ops.knn(X[s:e], X, k, q_row0=s, exclude_self=bool(self.zero_diag), metric=self.metric, prune=prune)
            P, rho, sigma = ops.umap_affinity_rows(dist, self.max_iter)
        self.rho_, self.eps_ = rho, sigma
        self.knn_ = (dist, idx)
        if timing:
            torch.cuda.synchronize()
            self.timings_ = {"knn+sigma": time.perf_counter() - t0}
        if not self.symmetrize:
            self.csr_ = None
            return P, idx
        self.logger.info("Symmetrizing affinity matrix...")
        ext = None
        if self.is_multi_gpu:
            counts, er, ec, ev = ops.symmetrize_export(P, idx, s, n, self.world_size, self.rank)
            ext = exchange_edges(counts, er, ec, ev)
        self.csr_ = ops.symmetrize_csr(P, idx, s, n, ext=ext, transpose_local=True)
        return self.csr_

    def __call__(self, X, return_indices=True, **kwargs):
        """``affinity/base.py:407-431`` contract."""
        out = self.compute_csr(X)
        if not self.symmetrize:
            P, idx = out
            return (P, idx) if return_indices else P
        values, indices = ops.csr_to_ell(*out)
        return (values, indices) if return_indices else values


class EntropicAffinity(_SparseAffinityBase):
    """``torchdr/affinity/entropic.py:118-312`` (sparse log form)."""

    def __init__(self, perplexity=30, max_iter=1000, sparsity=True, metric="sqeuclidean", zero_diag=True,
                 device="auto", backend=None, verbose=False, compile=False, distributed="auto",
                 _pre_processed=False, knn_order="auto"):
        self.perplexity = perplexity
        self.max_iter = max_iter
        super().__init__(metric=metric, zero_diag=zero_diag, device=device, backend=backend, verbose=verbose,
                         compile=compile, sparsity=sparsity, distributed=distributed, _pre_processed=_pre_processed,
                         knn_order=knn_order)

    def _compute_sparse_log_affinity(self, X):
        X = self._prepare(X)
        n = X.shape[0]
        perp = check_neighbor_param(self.perplexity, n)  # entropic.py:257
        if not self.sparsity:
            # dense N x N route (entropic.py:266-268, BASELINE config 3): full distance matrix, then one CTA
            # per row streams it once per bisection step; log_P overwrites C in place (4 N^2 bytes)
            C = ops.pairwise_full(X, None, metric=self.metric, exclude_diag=bool(self.zero_diag))
            target = float(torch.log(torch.tensor(perp)) + 1)
            log_n = float(torch.log(torch.tensor(float(n), dtype=torch.float32)))
            logP, eps, log_norm = ops.entropic_dense_rows(C, target, log_n, entropic_bound_scalars(n, perp),
                                                          self.max_iter, inplace=True)
            self.eps_ = eps
            self.log_normalization_ = log_norm.unsqueeze(1)
            return logP, None
        k = check_neighbor_param(3 * perp, n)  # entropic.py:259-265
        if k > _lib.TDR_MAX_K:
            raise NotImplementedError(f"[TorchDR-B200] 3*perplexity={k} exceeds the engine limit {_lib.TDR_MAX_K}.")
        if self.verbose:
            self.logger.info(f"Sparsity mode enabled, computing {k} nearest neighbors...")
        s, e = self._chunk(n)
        perm, prune = self._search_order(X)
        if perm is not None:
            from .reorder import unpermute_rows

            Xp = X[perm].contiguous()
            C, idx = ops.knn(Xp, Xp, k, q_row0=0, exclude_self=bool(self.zero_diag), metric=self.metric, prune=prune,
                             labels=perm.to(torch.int32))
            idx, C = unpermute_rows(perm, idx, C)
            del Xp
        else:
            C, idx = ops.knn(X[s:e], X, k, q_row0=s, exclude_self=bool(self.zero_diag), metric=self.metric, prune=prune)
        self.knn_ = (C, idx)
        target = float(torch.log(torch.tensor(perp)) + 1)  # entropic.py:272
        log_n = float(torch.log(torch.tensor(float(n), dtype=torch.float32)))  # entropic.py:308-310
        # entropic.py:280-287: the bracket is skipped on the multi-GPU path; tN = C_.shape[0]
        bounds = None if self.is_multi_gpu else entropic_bound_scalars(C.shape[0], perp)
        logP, eps, log_norm = ops.entropic_affinity_rows(C, target, log_n, bounds, self.max_iter)
        self.eps_ = eps
        self.log_normalization_ = log_norm.unsqueeze(1)
        return logP, idx

    def __call__(self, X, log=False, return_indices=True, **kwargs):
        """``affinity/base.py:522-561`` contract."""
        logP, idx = self._compute_sparse_log_affinity(X)
        out = logP if log else logP.exp_()
        return (out, idx) if return_indices else out


def log2_neighbors(k):
    return math.log2(k)
