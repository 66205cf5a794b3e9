#!/usr/bin/env python
"""Golden run of BASELINE.json configs[0] from the REAL reference (build container only):

    TSNE(perplexity=30, backend=None, device="cpu") on make_blobs(2000, 50, centers=10, random_state=0)

(SURVEY.md section 8d, row C1).  The run uses the estimator's defaults (early exaggeration 12 for 250 iterations,
lr "auto", momentum 0.5) with an injected normal initialisation; it is stopped after 50 iterations by max_iter —
t-SNE's schedule does not depend on max_iter — and the embedding is captured at 1, 2, 5, 10, 20, 50.  Only X, the
initialisation and the checkpoints are stored: the oracle recomputes the kNN graph and the entropic affinity, so the
fixture pins the whole CPU path of the reference at this size, not just the loop.

    python tests/golden/make_golden_c1.py
"""

import numpy as np
import torch
from make_golden import _capture_mixin, _import_reference, save


def main():
    _import_reference()
    from sklearn.datasets import make_blobs
    from torchdr import TSNE

    torch.set_num_threads(8)
    X, _ = make_blobs(n_samples=2000, n_features=50, centers=10, random_state=0)
    X = torch.from_numpy(X.astype(np.float32))
    g = torch.Generator().manual_seed(2000)
    Zinit = torch.randn(2000, 2, generator=g)
    cps = (1, 2, 5, 10, 20, 50)
    Cap = _capture_mixin(TSNE, 0, cps)
    m = Cap(perplexity=30, max_iter=50, init=Zinit, backend=None, device="cpu", random_state=0,
            process_duplicates=False, min_grad_norm=0.0)
    m.fit_transform(X)
    cap = m._cap
    save("tsne_c1_n2000_d50_p30", X=X, Zinit=Zinit, Z0=cap["Z0"], lr=np.asarray(cap["lr"]),
         eps_head=m.affinity_in.eps_[:64] if hasattr(m.affinity_in, "eps_") else np.zeros(0),
         P_head=cap["aff_vals"][:8], I_head=cap["aff_idx"][:8].to(torch.int32),
         **{f"Z_{s}": cap["Z"][s] for s in cps})


if __name__ == "__main__":
    main()
