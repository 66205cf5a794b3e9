"""Consumers of the kNN kernel outside the fit path.

``neighborhood_preservation`` (``torchdr/eval/neighborhood_preservation.py:20-186``) is the quality score the
reference uses to compare long runs (``benchmarks/umap_vs_largevis_distributed.py:97-107``), here also the
long-run parity criterion: a chaotic optimiser cannot be compared coordinate by coordinate after hundreds
of steps, but the neighbourhood structure of the result can.  ``knn_label_accuracy``
(``torchdr/eval/knn_labels.py:17-196``) is its label-based companion.  Both keep the reference's
signature; ``backend`` is accepted and ignored (there is one engine), and in distributed mode each rank
returns the result of its own row chunk, as the reference does.
"""

import numpy as np
import torch
import torch.distributed as dist

from .distance import _to_device_tensor, pairwise_distances
from .distributed import DistributedContext


def _resolve(distributed, device, X):
    distributed = dist.is_initialized() if distributed == "auto" else bool(distributed)
    if distributed:
        if not dist.is_initialized():
            raise RuntimeError(
                "[TorchDR] distributed=True requires launching with torchrun. "
                "Example: torchrun --nproc_per_node=4 your_script.py"
            )
        if device == "cpu":
            raise ValueError("[TorchDR] Distributed mode requires GPU (device cannot be 'cpu')")
        ctx = DistributedContext()
        return ctx, torch.device(f"cuda:{ctx.local_rank}")
    if device is None:
        device = X.device if isinstance(X, torch.Tensor) and X.is_cuda else "auto"
    return None, device


def _finish(values, return_per_sample, input_is_numpy):
    if return_per_sample:
        return values.detach().cpu().numpy() if input_is_numpy else values
    result = values.mean()
    return result.detach().cpu().numpy().item() if input_is_numpy else result


def neighborhood_preservation(X, Z, K=10, metric="sqeuclidean", backend=None, distributed="auto",
                              return_per_sample=False, device=None):
    """Mean fraction of each point's K nearest neighbours in X that are also among its K nearest in Z."""
    input_is_numpy = not isinstance(X, torch.Tensor) or not isinstance(Z, torch.Tensor)
    if len(X) != len(Z):
        raise ValueError(f"X and Z must have same number of samples, got {len(X)} and {len(Z)}")
    if K >= len(X):
        raise ValueError(f"K ({K}) must be less than number of samples ({len(X)})")
    ctx, device = _resolve(distributed, device, X)
    Xd = _to_device_tensor(X, device)
    Zd = _to_device_tensor(Z, Xd.device)
    _, nx = pairwise_distances(Xd, metric=metric, k=K, exclude_diag=True, return_indices=True, distributed_ctx=ctx)
    _, nz = pairwise_distances(Zd, metric=metric, k=K, exclude_diag=True, return_indices=True, distributed_ctx=ctx)
    matches = (nx.unsqueeze(2) == nz.unsqueeze(1)).any(dim=2)  # (n, K): neighborhood_preservation.py:166-171
    overlaps = matches.float().sum(dim=1) / K
    return _finish(overlaps, return_per_sample, input_is_numpy)


def knn_label_accuracy(X, labels, k=10, metric="euclidean", backend=None, exclude_self=True, distributed="auto",
                       return_per_sample=False, device=None):
    """Mean fraction of each point's k nearest neighbours that carry the point's own label."""
    if k < 1:
        raise ValueError(f"k must be at least 1, got {k}")
    input_is_numpy = not isinstance(X, torch.Tensor) or not isinstance(labels, torch.Tensor)
    if len(X) != len(labels):
        raise ValueError(f"X and labels must have same number of samples, got {len(X)} and {len(labels)}")
    n = len(X)
    if k >= n:
        raise ValueError(f"k ({k}) must be less than number of samples ({n})")
    ctx, device = _resolve(distributed, device, X)
    Xd = _to_device_tensor(X, device)
    lab = torch.as_tensor(np.asarray(labels) if not isinstance(labels, torch.Tensor) else labels).to(Xd.device)
    _, idx = pairwise_distances(Xd, metric=metric, k=k, exclude_diag=exclude_self, return_indices=True,
                                distributed_ctx=ctx)
    neighbor_labels = lab[idx.long()]
    if ctx is not None and ctx.is_initialized:
        s, e = ctx.compute_chunk_bounds(n)
        query = lab[s:e].unsqueeze(1)
    else:
        query = lab.unsqueeze(1)
    accuracies = (neighbor_labels == query).float().mean(dim=1)  # knn_labels.py:176-186
    return _finish(accuracies, return_per_sample, input_is_numpy)
