#!/usr/bin/env python
"""Generate golden fixtures from the REAL reference (TorchDR at /root/reference).

Run in the build container only (the GPU box has no /root/reference):

    python tests/golden/make_golden.py

The reference ships no expected arrays for this path, so
the fixtures are outputs of the reference itself on seeded inputs:
``backend=None, device="cpu"``, fp32.  An empty ``matplotlib`` stub package is
put on ``sys.path`` because ``torchdr/utils/visu.py:9`` imports it.

Negative samples: the reference draws them with ``torch.randint`` from the
global generator (``neighbor_embedding/base.py:629``).  To make the tables
reproducible without storing them, the subclass below wraps the *reference's
own* ``on_training_step_start`` and, for the duration of that call, routes
``torch.randint`` through a ``torch.Generator`` seeded with
``neg_seed(seed, step)``; tests rebuild the identical tables from the seed
(``tests/helpers.py:negative_table``).
"""

import os
import sys
import tempfile

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get("TORCHDR_REFERENCE", "/root/reference")


def _import_reference():
    stub = tempfile.mkdtemp(prefix="mplstub_")
    os.makedirs(os.path.join(stub, "matplotlib"))
    for f in ("__init__.py", "pylab.py", "pyplot.py"):
        open(os.path.join(stub, "matplotlib", f), "w").close()
    sys.path.insert(0, stub)
    sys.path.insert(0, REF)
    import torchdr  # noqa: F401

    return torchdr


def neg_seed(seed, step):
    return seed * 1000003 + step


def blobs(n, d, centers, seed, spread=1.0, scale=6.0):
    g = torch.Generator().manual_seed(seed)
    c = torch.randn(centers, d, generator=g) * scale
    lab = torch.randint(0, centers, (n,), generator=g)
    return (c[lab] + torch.randn(n, d, generator=g) * spread).float().contiguous()


def _capture_mixin(base, seed, checkpoints):
    class Cap(base):
        def on_affinity_computation_end(self):
            self._cap = {"aff_vals": self.affinity_in_.detach().clone(),
                         "aff_idx": self.NN_indices_.detach().clone()}
            super().on_affinity_computation_end()
            if hasattr(self, "epochs_per_sample"):
                self._cap["eps_per_sample"] = self.epochs_per_sample.detach().clone()
            self._cap["Z"] = {}
            self._cap["lr"] = []
            self._cap["grad"] = {}

        def on_training_step_start(self):
            step = int(self.n_iter_)
            if step == 0:
                self._cap["Z0"] = self.embedding_.detach().clone()
            self._cap["lr"].append(float(self.optimizer_.param_groups[0]["lr"]))
            g = torch.Generator().manual_seed(neg_seed(seed, step))
            orig = torch.randint

            def seeded(*a, **kw):
                kw.pop("device", None)
                return orig(*a, generator=g, **kw)

            torch.randint = seeded
            try:
                super().on_training_step_start()
            finally:
                torch.randint = orig
            if step == 0 and hasattr(self, "neg_indices_"):
                self._cap["neg0"] = self.neg_indices_.detach().clone()

        def on_training_step_end(self):
            step = int(self.n_iter_) + 1
            if step in checkpoints:
                self._cap["Z"][step] = self.embedding_.detach().clone()
                if self.embedding_.grad is not None:
                    self._cap["grad"][step] = self.embedding_.grad.detach().clone()
            super().on_training_step_end()

        def clear_memory(self):  # keep captured state
            cap = self._cap
            super().clear_memory()
            self._cap = cap

    return Cap


def save(name, **arrs):
    out = {}
    for k, v in arrs.items():
        if isinstance(v, torch.Tensor):
            v = v.detach().cpu().numpy()
        out[k] = np.asarray(v)
    path = os.path.join(HERE, name + ".npz")
    np.savez_compressed(path, **out)
    print(f"{name}: {os.path.getsize(path) / 1024:.1f} KiB", {k: v.shape for k, v in out.items()})


def main():
    torchdr = _import_reference()
    from torchdr.distance import pairwise_distances
    from torchdr.affinity import UMAPAffinity, EntropicAffinity
    from torchdr.affinity.entropic import _bounds_entropic_affinity
    from torchdr.utils.sparse import symmetrize_sparse
    from torchdr import UMAP, LargeVis, TSNE

    torch.set_num_threads(8)
    meta = {"torch": torch.__version__, "numpy": np.__version__}

    # ---------------- kNN -------------------------------------------------
    for tag, (n, d, k, seed) in {"knn_n300_d16_k15": (300, 16, 15, 1),
                                 "knn_n2000_d50_k90": (2000, 50, 90, 2),
                                 "knn_n1500_d128_k15": (1500, 128, 15, 3)}.items():
        X = blobs(n, d, 10, seed)
        C, I = pairwise_distances(X, metric="sqeuclidean", backend=None, exclude_diag=True,
                                  k=k, return_indices=True)
        Ce, Ie = pairwise_distances(X, metric="euclidean", backend=None, exclude_diag=True,
                                    k=k, return_indices=True)
        save(tag, X=X, C=C, I=I.to(torch.int32), Ce=Ce, Ie=Ie.to(torch.int32), k=k)

    # full matrix + k >= n behaviour
    X = blobs(64, 8, 3, 4)
    Cf = pairwise_distances(X, metric="sqeuclidean", backend=None, exclude_diag=True)
    Cf2, If2 = pairwise_distances(X, metric="sqeuclidean", backend=None, exclude_diag=False,
                                  k=70, return_indices=True)
    assert If2 is None
    Y = blobs(40, 8, 3, 5)
    Cxy, Ixy = pairwise_distances(X, Y, metric="sqeuclidean", backend=None, k=5, return_indices=True)
    save("pairwise_full_n64", X=X, Y=Y, C_excl=Cf, C_kge_n=Cf2, Cxy=Cxy, Ixy=Ixy.to(torch.int32))

    # ---------------- UMAP affinity + symmetrise + loop -------------------
    n, d, k, seed, T = 300, 16, 15, 1, 100
    X = blobs(n, d, 10, seed)
    aff = UMAPAffinity(n_neighbors=k, max_iter=100, backend=None, device="cpu", symmetrize=False)
    P, I = aff(X, return_indices=True)
    rho, sig = aff.rho_.clone(), aff.eps_.clone()
    Vs, Is = symmetrize_sparse(P, I, mode="sum_minus_prod")
    g = torch.Generator().manual_seed(100 + seed)
    Zinit = torch.randn(n, 2, generator=g)
    checkpoints = (1, 2, 5, 20, 50, 100)
    Cap = _capture_mixin(UMAP, seed, checkpoints)
    m = Cap(n_neighbors=k, max_iter=T, init=Zinit, backend=None, device="cpu",
            random_state=0, process_duplicates=False, min_grad_norm=0.0)
    Zfinal = m.fit_transform(X)
    cap = m._cap
    assert torch.equal(cap["aff_idx"], Is) and torch.equal(cap["aff_vals"], Vs)
    save("umap_n300_d16_k15", X=X, P=P, I=I.to(torch.int32), rho=rho, sigma=sig,
         sym_vals=Vs, sym_idx=Is.to(torch.int32), eps_per_sample=cap["eps_per_sample"],
         Zinit=Zinit, Z0=cap["Z0"], lr=np.asarray(cap["lr"], dtype=np.float64),
         neg0=cap["neg0"].to(torch.int32), a=m._a, b=m._b, max_iter=T, seed=seed,
         Zfinal=Zfinal.detach(),
         **{f"Z_{s}": cap["Z"][s] for s in checkpoints})

    # ---------------- Entropic affinity -----------------------------------
    n, d, perp, seed = 300, 16, 10, 6
    X = blobs(n, d, 10, seed)
    ea = EntropicAffinity(perplexity=perp, max_iter=100, backend=None, device="cpu")
    logP, I = ea(X, log=True, return_indices=True)
    C, I2 = pairwise_distances(X, metric="sqeuclidean", backend=None, exclude_diag=True,
                               k=3 * perp, return_indices=True)
    assert torch.equal(I, I2)
    b0, b1 = _bounds_entropic_affinity(C, torch.tensor(perp), device="cpu", dtype=torch.float32)
    ea2 = EntropicAffinity(perplexity=perp, max_iter=100, backend=None, device="cpu")
    ea2.is_multi_gpu = True  # entropic.py:280-282 — default bracket begin=end=1
    logP_nb, _ = ea2(X, log=True, return_indices=True)
    save("entropic_n300_d16_p10", X=X, C=C, I=I.to(torch.int32), logP=logP, eps=ea.eps_,
         log_norm=ea.log_normalization_.squeeze(-1), begin=b0, end=b1,
         logP_nobounds=logP_nb, eps_nobounds=ea2.eps_, perplexity=perp)

    # dense route (sparsity=False, entropic.py:266-268): full N x N log-affinity
    ead = EntropicAffinity(perplexity=perp, max_iter=100, backend=None, device="cpu", sparsity=False)
    logPd = ead(X, log=True, return_indices=False)
    save("entropic_dense_n300_d16_p10", X=X, logP=logPd, eps=ead.eps_, log_norm=ead.log_normalization_.squeeze(-1),
         perplexity=perp)

    # perplexity 30 on the C1-like data (k=90)
    X = blobs(2000, 50, 10, 2)
    ea = EntropicAffinity(perplexity=30, max_iter=100, backend=None, device="cpu")
    logP, I = ea(X, log=True, return_indices=True)
    save("entropic_n2000_d50_p30", eps=ea.eps_, log_norm=ea.log_normalization_.squeeze(-1),
         logP_head=logP[:64], I_head=I[:64].to(torch.int32))

    # ---------------- LargeVis / TSNE loops -------------------------------
    n, d, perp, seed = 300, 16, 10, 7
    X = blobs(n, d, 10, seed)
    g = torch.Generator().manual_seed(100 + seed)
    Zinit = torch.randn(n, 2, generator=g)
    cps = (1, 2, 5, 10, 30)
    Cap = _capture_mixin(LargeVis, seed, cps)
    m = Cap(perplexity=perp, max_iter=30, init=Zinit, backend=None, device="cpu",
            random_state=0, process_duplicates=False, min_grad_norm=0.0)
    m.fit_transform(X)
    cap = m._cap
    save("largevis_n300_d16_p10", X=X, P=cap["aff_vals"], I=cap["aff_idx"].to(torch.int32),
         Zinit=Zinit, Z0=cap["Z0"], lr=np.asarray(cap["lr"]), seed=seed,
         neg0=cap["neg0"].to(torch.int32),
         **{f"Z_{s}": cap["Z"][s] for s in cps}, **{f"G_{s}": cap["grad"][s] for s in cps})

    cps = (1, 2, 5, 10, 11, 12, 20)
    Cap = _capture_mixin(TSNE, seed, cps)
    m = Cap(perplexity=perp, max_iter=20, init=Zinit, backend=None, device="cpu",
            early_exaggeration_iter=10, random_state=0, process_duplicates=False,
            min_grad_norm=0.0)
    m.fit_transform(X)
    cap = m._cap
    save("tsne_n300_d16_p10", X=X, P=cap["aff_vals"], I=cap["aff_idx"].to(torch.int32),
         Zinit=Zinit, Z0=cap["Z0"], lr=np.asarray(cap["lr"]), exag_iter=10,
         **{f"Z_{s}": cap["Z"][s] for s in cps}, **{f"G_{s}": cap["grad"][s] for s in cps})

    # ---------------- schedules ------------------------------------------
    save("meta", **{k: np.asarray(v) for k, v in meta.items()})


if __name__ == "__main__":
    main()
