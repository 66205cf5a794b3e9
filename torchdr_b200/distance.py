"""Distance seam: ``pairwise_distances`` with the reference's signature and return contract.

Mirrors ``torchdr/distance/base.py:22-249`` (dispatcher) over the torch-backend semantics of
``torchdr/distance/torch.py:21-125``; the compute is ``tdr_knn_f32`` / ``tdr_pairwise_full_f32``.
There is a single backend (sm_100a CUDA): ``backend`` is accepted for signature compatibility and
ignored; CPU tensors are moved to the current CUDA device and the output stays there
(``distance/torch.py:119-125``: "output stays on computation device").
"""

import numpy as np
import torch

from . import _lib, ops

LIST_METRICS_B200 = ["euclidean", "sqeuclidean"]
_REFERENCE_ONLY = ["manhattan", "angular", "sqhyperbolic"]  # distance/torch.py:12-18, not on the hot path


def _to_device_tensor(X, device="auto"):
    if isinstance(X, np.ndarray):
        X = torch.from_numpy(X)
    if not isinstance(X, torch.Tensor):
        raise TypeError("[TorchDR] ERROR : input must be a torch.Tensor or numpy array.")
    if device not in (None, "auto"):
        tgt = torch.device(device)
        if tgt.type != "cuda":
            raise _lib.B200EngineError(f"[TorchDR-B200] device={device!r}: the engine has no CPU path.")
        X = X.to(tgt)
    elif X.device.type != "cuda":
        if not torch.cuda.is_available():
            raise _lib.B200EngineError("[TorchDR-B200] no CUDA device visible: the engine has no CPU path.")
        X = X.to(torch.device("cuda", torch.cuda.current_device()))
    return X


def _check_metric(metric):
    if metric in LIST_METRICS_B200:
        return
    if metric in _REFERENCE_ONLY:
        raise NotImplementedError(
            f"[TorchDR-B200] the '{metric}' distance is outside the accelerated path (sqeuclidean / euclidean)."
        )
    raise ValueError(f"[TorchDR] ERROR : The '{metric}' distance is not supported.")  # distance/torch.py:63-64


def pairwise_distances(X, Y=None, metric="euclidean", backend=None, exclude_diag=False, k=None,
                       return_indices=False, device="auto", distributed_ctx=None):
    """``torchdr/distance/base.py:22-32`` signature.

    Returns ``C`` or ``(C, indices)``: with ``k`` the k smallest distances per row in ascending
    order and int32 indices; ``k >= n_db`` returns the full matrix and ``None``
    (``utils/utils.py:203-204``); without ``k`` the dense matrix.  With ``distributed_ctx`` the
    queries are this rank's row chunk against the full database (``base.py:160-211``) and the
    row's own index is excluded by id (the 1e12-diagonal rule of ``torch.py:111-116``; the
    reference's k+1-then-drop-column-0 shortcut, ``base.py:191-206``, mishandles duplicates).
    """
    _check_metric(metric)
    X = _to_device_tensor(X, device)
    if X.dtype != torch.float32:
        X = X.float()
    same = Y is None or Y is X
    if not same:
        Y = _to_device_tensor(Y, device).to(X.device)
        if Y.dtype != torch.float32:
            Y = Y.float()

    if distributed_ctx is not None and distributed_ctx.is_initialized:
        if k is None:
            raise ValueError(
                "[TorchDR] Distributed mode requires sparse computation with k-NN. "
                "k cannot be None when distributed_ctx is provided."
            )
        if not same:
            raise ValueError(
                "[TorchDR] Distributed mode does not support cross-distance computation. "
                "Y must be None when distributed_ctx is provided."
            )
        start, end = distributed_ctx.compute_chunk_bounds(X.shape[0])
        C, idx = ops.knn(X[start:end], X, int(k), q_row0=start, exclude_self=bool(exclude_diag), metric=metric)
        return (C, idx) if return_indices else C

    db = X if same else Y
    if k is not None and int(k) < db.shape[0]:
        k = int(k)
        if k > _lib.TDR_MAX_K:
            raise NotImplementedError(f"[TorchDR-B200] k={k} exceeds the engine limit of {_lib.TDR_MAX_K}.")
        avail = db.shape[0] - (1 if (same and exclude_diag) else 0)
        if k > avail:
            # the reference would return the 1e12 self-distance as last neighbour; nothing sensible to mirror
            raise ValueError(f"[TorchDR] ERROR : k={k} exceeds the {avail} available neighbours.")
        C, idx = ops.knn(X, db, k, q_row0=0, exclude_self=bool(same and exclude_diag), metric=metric)
    else:
        C = ops.pairwise_full(X, None if same else Y, metric=metric, exclude_diag=bool(same and exclude_diag))
        idx = None
    return (C, idx) if return_indices else C


def pairwise_distances_indexed(X, query_indices=None, key_indices=None, Y=None, metric="sqeuclidean", backend=None,
                               device="auto"):
    """``torchdr/distance/base.py:252-405`` for the case the neighbor-embedding losses use: per-query key lists
    (2-D ``key_indices``), optional 1-D ``query_indices``; returns ``[n_queries, n_keys]`` in the exact-difference
    form (``base.py:384-385``).  Other index shapes are outside the accelerated path."""
    _check_metric(metric)
    X = _to_device_tensor(X, device)
    if X.dtype != torch.float32:
        X = X.float()
    Yd = X if Y is None else _to_device_tensor(Y, device).to(X.device).float()
    if key_indices is None or key_indices.dim() != 2:
        raise NotImplementedError("[TorchDR-B200] pairwise_distances_indexed needs 2-D key_indices (per-query keys).")
    if query_indices is not None and query_indices.dim() != 1:
        raise NotImplementedError("2D query indices not yet supported")  # base.py:340
    key = key_indices.to(X.device)
    if key.dtype not in (torch.int32, torch.int64):
        key = key.long()
    q = None if query_indices is None else query_indices.to(X.device)
    n_q = X.shape[0] if q is None else q.numel()
    assert key.shape[0] == n_q, f"key_indices first dim {key.shape[0]} must match number of queries {n_q}"  # base.py:349-352
    return ops.indexed_distances(X, key, Y=Yd, query_idx=q, metric=metric)
