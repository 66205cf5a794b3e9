#!/usr/bin/env python
"""Golden fixtures for the InfoTSNE and SNE loops, from the REAL reference (build container only).

    python tests/golden/make_golden_more.py

Same conventions as ``make_golden.py`` (whose helpers it reuses): ``backend=None, device="cpu"``,
fp32, injected initial embedding, negatives drawn through a generator seeded with
``neg_seed(seed, step)`` so tests can rebuild the tables (``tests/helpers.py:negative_table``).
"""

import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from make_golden import _capture_mixin, _import_reference, blobs, save  # noqa: E402


def main():
    _import_reference()
    from torchdr import SNE, InfoTSNE

    torch.set_num_threads(8)
    n, d, perp, seed = 300, 16, 10, 7
    X = blobs(n, d, 10, seed)
    g = torch.Generator().manual_seed(100 + seed)
    Zinit = torch.randn(n, 2, generator=g)

    cps = (1, 2, 5, 10, 11, 12, 20)
    Cap = _capture_mixin(InfoTSNE, seed, cps)
    m = Cap(perplexity=perp, max_iter=20, init=Zinit, backend=None, device="cpu", n_negatives=50,
            early_exaggeration_iter=10, random_state=0, process_duplicates=False, min_grad_norm=0.0)
    m.fit_transform(X)
    cap = m._cap
    save("infotsne_n300_d16_p10", X=X, P=cap["aff_vals"], I=cap["aff_idx"].to(torch.int32),
         Zinit=Zinit, Z0=cap["Z0"], lr=np.asarray(cap["lr"]), seed=seed, exag_iter=10, n_neg=50,
         neg0=cap["neg0"].to(torch.int32),
         **{f"Z_{s}": cap["Z"][s] for s in cps}, **{f"G_{s}": cap["grad"][s] for s in cps})

    cps = (1, 2, 5, 10, 20)
    Cap = _capture_mixin(SNE, seed, cps)
    m = Cap(perplexity=perp, max_iter=20, init=Zinit, backend=None, device="cpu",
            random_state=0, process_duplicates=False, min_grad_norm=0.0)
    m.fit_transform(X)
    cap = m._cap
    save("sne_n300_d16_p10", X=X, P=cap["aff_vals"], I=cap["aff_idx"].to(torch.int32),
         Zinit=Zinit, Z0=cap["Z0"], lr=np.asarray(cap["lr"]),
         **{f"Z_{s}": cap["Z"][s] for s in cps}, **{f"G_{s}": cap["grad"][s] for s in cps})


if __name__ == "__main__":
    main()
