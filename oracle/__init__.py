"""CPU oracle for the neighbor-embedding hot path — TEST INFRASTRUCTURE ONLY.

This package is a CPU restatement (torch-CPU / numpy ops, the same substrate
the reference runs on with ``backend=None, device="cpu"``) of the one path this
repository accelerates:

    pairwise distance / exact kNN  ->  per-row sigma/rho bisection affinity
    ->  graph symmetrisation  ->  sampled attraction/repulsion + SGD update.

Every function cites the reference file:line it follows (paths relative to the
TorchDR tree, ``/root/reference`` in the build container).

Rules (enforced by ``tests/test_layout.py``):
  * only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s
    ``cpu_baseline`` / ``--impl reference`` legs may import this package;
  * nothing under ``torchdr_b200/`` imports it — the product path has no CPU
    fallback and fails loudly when the CUDA library is missing.

Parity pinning: the reference ships no golden vectors for this path (its tests
check properties only).  The oracle is therefore pinned
against outputs of the reference itself, generated in the build container by
``tests/golden/make_golden.py`` / ``make_golden_more.py`` (which import ``/root/reference/torchdr``) and
committed as ``tests/golden/*.npz``.  ``tests/test_oracle_golden.py`` checks
every oracle function against those files.
"""

from .knn import (  # noqa: F401
    knn_dense,
    knn_chunked,
    pairwise_full,
    knn_ambiguity,
    tile_prune_plan,
    knn_with_tile_mask,
)
from .root_search import bisect_rows  # noqa: F401
from .affinity import (  # noqa: F401
    umap_affinity_rows,
    entropic_affinity_rows,
    entropic_bounds,
    clamp_neighbor_param,
)
from .graph import symmetrize_ell, ell_to_csr, csr_to_ell  # noqa: F401
from .umap import (  # noqa: F401
    find_ab,
    umap_edge_schedule,
    umap_step,
    umap_run,
    adjust_negatives,
    linear_lr_sequence,
)
from .largevis import largevis_run  # noqa: F401
from .tsne import tsne_run  # noqa: F401
from .infotsne import infotsne_run, infotsne_loss  # noqa: F401
from .sne import sne_run, sne_loss  # noqa: F401
from .partition import chunk_bounds, owner_of  # noqa: F401
