"""Consumer of the kNN kernel outside the fit path: neighbourhood preservation
(``torchdr/eval/neighborhood_preservation.py:148-176``) — the quality score the reference uses to
compare long runs (``benchmarks/umap_vs_largevis_distributed.py:97-107``), here as the long-run
parity criterion: a chaotic optimiser cannot be compared coordinate by coordinate after hundreds of
steps, but the neighbourhood structure of the result can."""

import torch

from .distance import pairwise_distances


def neighborhood_preservation(X, Z, K=10, metric="sqeuclidean", return_per_sample=False):
    """Mean fraction of each point's K nearest neighbours in X that are also among its K nearest in Z."""
    _, nx = pairwise_distances(X, metric=metric, k=K, exclude_diag=True, return_indices=True)
    _, nz = pairwise_distances(Z, metric=metric, k=K, exclude_diag=True, return_indices=True)
    nz = nz.to(nx.device)
    matches = (nx.unsqueeze(2) == nz.unsqueeze(1)).any(dim=2)  # (n, K): same rule as the reference, lines 166-171
    overlaps = matches.float().sum(dim=1) / K
    return overlaps if return_per_sample else overlaps.mean()
