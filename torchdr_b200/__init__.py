"""B200-native neighbor-embedding engine behind TorchDR's distance / Affinity / NeighborEmbedding seams.

Host code is Python over PyTorch tensors; all compute is hand-written sm_100a CUDA reached
through the C ABI of ``include/tdrb200.h`` (``torchdr_b200/lib/libtdrb200.so``).
"""

from . import _lib  # noqa: F401
from .distance import pairwise_distances, pairwise_distances_indexed, LIST_METRICS_B200  # noqa: F401
from .affinity import UMAPAffinity, EntropicAffinity  # noqa: F401
from .neighbor_embedding import UMAP, LargeVis, TSNE, InfoTSNE, SNE  # noqa: F401
from .distributed import DistributedContext  # noqa: F401
from .eval import neighborhood_preservation, knn_label_accuracy  # noqa: F401

__all__ = [
    "pairwise_distances",
    "pairwise_distances_indexed",
    "UMAPAffinity",
    "EntropicAffinity",
    "UMAP",
    "LargeVis",
    "TSNE",
    "InfoTSNE",
    "SNE",
    "DistributedContext",
    "neighborhood_preservation",
    "knn_label_accuracy",
]
