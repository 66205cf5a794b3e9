"""The oracle against the REAL reference, live, on inputs beyond the committed fixtures.

Runs only where the reference tree is mounted (the build container, /root/reference); on the GPU box — which has no
reference — every test here is skipped, and nothing in the `-m gpu` suite, smoke() or bench.py reads that path.
Each case draws fresh seeded inputs (several shapes, duplicates, tiny n) and requires bit-identical outputs of the
reference's `backend=None, device="cpu"` functions and of their restatements in oracle/.
"""

import os
import sys
import tempfile

import numpy as np
import pytest
import torch

import oracle
from helpers import blobs

REF = os.environ.get("TORCHDR_REFERENCE", "/root/reference")
pytestmark = pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "torchdr")), reason="reference tree not mounted")


@pytest.fixture(scope="module")
def ref():
    stub = tempfile.mkdtemp(prefix="mplstub_")
    os.makedirs(os.path.join(stub, "matplotlib"))
    for f in ("__init__.py", "pylab.py", "pyplot.py"):  # torchdr/utils/visu.py:9 imports matplotlib unconditionally
        open(os.path.join(stub, "matplotlib", f), "w").close()
    sys.path.insert(0, stub)
    sys.path.insert(0, REF)
    try:
        import torchdr
    finally:
        sys.path.remove(REF)
        sys.path.remove(stub)
    torch.set_num_threads(min(8, torch.get_num_threads()))
    return torchdr


CASES = [(64, 3, 5, 0), (257, 17, 15, 1), (500, 50, 90, 2), (1000, 128, 15, 3), (130, 8, 30, 4)]


def _data(n, d, seed, duplicates=False):
    X = blobs(n, d, max(2, n // 60), seed)
    if duplicates:
        X[n // 2:n // 2 + 7] = X[:7]
    return X


@pytest.mark.parametrize("n,d,k,seed", CASES)
@pytest.mark.parametrize("metric", ["sqeuclidean", "euclidean"])
def test_knn_equals_reference(ref, n, d, k, seed, metric):
    from torchdr.distance import pairwise_distances

    X = _data(n, d, seed)
    C_ref, I_ref = pairwise_distances(X, metric=metric, backend=None, exclude_diag=True, k=k, return_indices=True)
    C, I = oracle.knn_dense(X, k, metric, True)
    assert torch.equal(I, I_ref) and torch.equal(C, C_ref)
    full_ref = pairwise_distances(X, metric=metric, backend=None)
    assert torch.equal(oracle.pairwise_full(X, None, metric), full_ref)


@pytest.mark.parametrize("n,d,k,seed", CASES)
@pytest.mark.parametrize("duplicates", [False, True])
def test_umap_affinity_equals_reference(ref, n, d, k, seed, duplicates):
    from torchdr.affinity import UMAPAffinity

    X = _data(n, d, seed, duplicates)
    kk = min(k, n - 2)
    aff = UMAPAffinity(n_neighbors=kk, max_iter=100, backend=None, device="cpu")
    V_ref, J_ref = aff(X, return_indices=True)
    C, I = oracle.knn_dense(X, kk)
    P, rho, sigma = oracle.umap_affinity_rows(C, kk, max_iter=100)
    assert torch.equal(rho, aff.rho_.squeeze(-1) if aff.rho_.dim() > 1 else aff.rho_)
    assert torch.equal(sigma, aff.eps_.squeeze(-1) if aff.eps_.dim() > 1 else aff.eps_)
    V, J = oracle.symmetrize_ell(P, I)
    assert torch.equal(J, J_ref.long()) and torch.equal(V, V_ref)
    # the CSR form the engine uses carries the same entries
    rp, col, val = oracle.ell_to_csr(V, J)
    V2, J2 = oracle.csr_to_ell(rp, col, val)
    assert torch.equal(V2, V) and torch.equal(J2, J)


@pytest.mark.parametrize("n,d,perp,seed", [(64, 3, 5, 0), (300, 16, 10, 1), (700, 50, 30, 2), (200, 128, 40, 3)])
def test_entropic_affinity_equals_reference(ref, n, d, perp, seed):
    from torchdr.affinity import EntropicAffinity

    X = _data(n, d, seed)
    aff = EntropicAffinity(perplexity=perp, max_iter=100, backend=None, device="cpu")
    logP_ref, I_ref = aff(X, log=True, return_indices=True)
    p = oracle.clamp_neighbor_param(perp, n)
    k = oracle.clamp_neighbor_param(3 * p, n)
    C, I = oracle.knn_dense(X, k)
    logP, eps, log_norm = oracle.entropic_affinity_rows(C, p, n_total=n, max_iter=100)
    assert torch.equal(I, I_ref) and torch.equal(logP, logP_ref)
    assert torch.equal(eps, aff.eps_.reshape(-1))


@pytest.mark.parametrize("seed", [0, 1, 2])
def test_umap_schedule_and_steps_equal_reference(ref, seed):
    """Edge schedule (umap.py:215-234) and three optimisation steps from a shared state: the reference's own
    _compute_gradients on its ELL arrays vs oracle.umap_step."""
    from torchdr import UMAP

    n, d, k, T = 200, 12, 10, 3
    X = _data(n, d, 10 + seed)
    g = torch.Generator().manual_seed(seed)
    Zinit = torch.randn(n, 2, generator=g)
    snaps, negs = {}, {}

    class Cap(UMAP):
        def on_training_step_start(self):
            super().on_training_step_start()
            negs[int(self.n_iter_)] = self.neg_indices_.clone()

        def on_training_step_end(self):
            snaps[int(self.n_iter_) + 1] = self.embedding_.detach().clone()
            super().on_training_step_end()

    m = Cap(n_neighbors=k, max_iter=T, init=Zinit, backend=None, device="cpu", random_state=seed,
            process_duplicates=False, min_grad_norm=0.0)
    m.fit_transform(X)
    C, I = oracle.knn_dense(X, k)
    P, _, _ = oracle.umap_affinity_rows(C, k, max_iter=100)
    V, J = oracle.symmetrize_ell(P, I)
    per, nxt = oracle.umap_edge_schedule(V, T)
    Z0 = 1e-4 * Zinit / Zinit[:, 0].std()
    a, b = oracle.find_ab()
    lrs = oracle.linear_lr_sequence(1.0, T, T)
    Z, nx = Z0, nxt
    for t_ in range(T):
        Z, nx = oracle.umap_run(Z, J, per, nx, [negs[t_]], [lrs[t_]], a, b, n_iter0=t_)
        assert torch.equal(Z, snaps[t_ + 1]), t_


@pytest.mark.parametrize("name,seed", [("LargeVis", 0), ("LargeVis", 1), ("TSNE", 0), ("TSNE", 1), ("InfoTSNE", 0), ("SNE", 0)])
def test_momentum_estimators_equal_reference(ref, name, seed):
    """Four steps of the autograd-mode estimators (below ATen's parallel grain size, where the reference is itself
    bit-reproducible), across the early-exaggeration switch at step 2, with the negatives the reference drew."""
    import torchdr

    n, d, perp, T = 150, 10, 8, 4
    X = _data(n, d, 20 + seed)
    Zinit = torch.randn(n, 2, generator=torch.Generator().manual_seed(seed))
    snaps, negs, cap = {}, {}, {}

    class Cap(getattr(torchdr, name)):
        def on_affinity_computation_end(self):
            cap["P"], cap["I"] = self.affinity_in_.detach().clone(), self.NN_indices_.detach().clone()
            super().on_affinity_computation_end()

        def on_training_step_start(self):
            super().on_training_step_start()
            if hasattr(self, "neg_indices_"):
                negs[int(self.n_iter_)] = self.neg_indices_.clone()

        def on_training_step_end(self):
            snaps[int(self.n_iter_) + 1] = self.embedding_.detach().clone()
            super().on_training_step_end()

    kw = {"early_exaggeration_iter": 2} if name in ("TSNE", "InfoTSNE") else {}
    if name == "InfoTSNE":
        kw["n_negatives"] = 20
    m = Cap(perplexity=perp, max_iter=T, init=Zinit, backend=None, device="cpu", random_state=seed,
            process_duplicates=False, min_grad_norm=0.0, **kw)
    m.fit_transform(X)
    p = oracle.clamp_neighbor_param(perp, n)
    C, I = oracle.knn_dense(X, oracle.clamp_neighbor_param(3 * p, n))
    P = oracle.entropic_affinity_rows(C, p, n_total=n, max_iter=100)[0].exp()
    assert torch.equal(P, cap["P"]) and torch.equal(I, cap["I"])
    Z0 = 1e-4 * Zinit / Zinit[:, 0].std()
    for T_ in range(1, T + 1):
        if name == "LargeVis":
            Z = oracle.largevis_run(Z0, P, I, [negs[t_] for t_ in range(T_)], T_)[0]
        elif name == "TSNE":
            Z = oracle.tsne_run(Z0, P, I, T_, exag_iter=2)
        elif name == "InfoTSNE":
            Z = oracle.infotsne_run(Z0, P, I, [negs[t_] for t_ in range(T_)], T_, exag_iter=2)[0]
        else:
            Z = oracle.sne_run(Z0, P, I, T_)
        assert torch.equal(Z, snaps[T_]), (name, T_)


@pytest.mark.parametrize("name", ["UMAP", "LargeVis"])
def test_discard_nns_host_flow_equals_reference(ref, name, monkeypatch):
    """discard_NNs=True: the engine's host flow (public estimator on the CPU stand-ins of tests/fake_ops.py) against the
    reference, same seed: the exclusion table (NE base.py:578-615), the negatives of every step (:638-647, drawn from
    the global generator right after seeding, as the reference does) and the embedding after 3 steps are identical."""
    import fake_ops
    import torchdr

    import torchdr_b200 as tb

    n, d, T = 160, 8, 3
    X = _data(n, d, 31)
    Zinit = torch.randn(n, 2, generator=torch.Generator().manual_seed(5))
    kw = {"n_neighbors": 8} if name == "UMAP" else {"perplexity": 6}
    seen = {"ref": {}, "eng": {}}

    def capture(base, tag):
        class Cap(base):
            def on_training_step_start(self):
                super().on_training_step_start()
                if int(self.n_iter_) == 0:
                    seen[tag]["excl"] = self.negative_exclusion_indices_.clone()
                seen[tag][f"neg{int(self.n_iter_)}"] = self.neg_indices_.clone()

        return Cap

    mr = capture(getattr(torchdr, name), "ref")(max_iter=T, init=Zinit, backend=None, device="cpu", random_state=3,
                                               process_duplicates=False, min_grad_norm=0.0, discard_NNs=True, **kw)
    Zr = mr.fit_transform(X)
    fake_ops.install(monkeypatch)
    fake_ops.install_entropic(monkeypatch)
    me = capture(getattr(tb, name), "eng")(max_iter=T, init=Zinit, random_state=3, process_duplicates=False,
                                          min_grad_norm=0.0, discard_NNs=True, **kw)
    Ze = me.fit_transform(X)
    assert torch.equal(seen["eng"]["excl"], seen["ref"]["excl"])
    for t_ in range(T):
        assert torch.equal(seen["eng"][f"neg{t_}"], seen["ref"][f"neg{t_}"]), t_
    assert torch.equal(Ze, Zr)


@pytest.mark.parametrize("name,opt,okw,lr", [
    ("TSNE", "Adam", None, 0.5),
    ("LargeVis", "SGD", {"momentum": 0.9, "nesterov": True}, 20.0),
    ("UMAP", "Adam", None, 0.05),
    ("UMAP", "SGD", {"momentum": 0.7}, 0.5),
    ("SNE", "RMSprop", {"alpha": 0.9}, 0.1),
])
def test_any_torch_optimizer_host_flow_equals_reference(ref, name, opt, okw, lr, monkeypatch):
    """`optimizer=` / `optimizer_kwargs=` beyond the fused SGD(+momentum) kernels: the engine hands the gradient its
    kernels computed to the reference's own optimiser object (NE base.py:312-343, affinity_matcher.py:395-429).  Host
    flow on the CPU stand-ins vs the live reference, same seed and initialisation: identical embeddings after 4 steps
    (the TSNE case crosses the early-exaggeration rebuild at step 2)."""
    import fake_ops
    import torchdr

    import torchdr_b200 as tb

    n, d, T = 150, 9, 4
    X = _data(n, d, 41)
    Zinit = torch.randn(n, 2, generator=torch.Generator().manual_seed(6))
    kw = {"n_neighbors": 8} if name == "UMAP" else {"perplexity": 6}
    if name == "TSNE":
        kw["early_exaggeration_iter"] = 2
    common = dict(max_iter=T, init=Zinit, random_state=4, process_duplicates=False, min_grad_norm=0.0, optimizer=opt,
                  optimizer_kwargs=okw, lr=lr, **kw)
    Zr = getattr(torchdr, name)(backend=None, device="cpu", **common).fit_transform(X)
    fake_ops.install(monkeypatch)
    fake_ops.install_entropic(monkeypatch)
    # the stand-in for the fused UMAP loop draws its own negatives; the reference draws from the global generator:
    # route the engine through the hook path with the reference's draw (NE base.py:629-636)
    cls = getattr(tb, name)
    if name == "UMAP":
        class cls(tb.UMAP):  # noqa: F811
            def on_training_step_start(self):
                raw = torch.randint(0, self.n_samples_in_ - 1, (n, self.n_negatives))
                self.neg_indices_ = raw + (raw >= torch.arange(n).unsqueeze(1)).long()
    elif name == "LargeVis":
        class cls(tb.LargeVis):  # noqa: F811
            def on_training_step_start(self):
                raw = torch.randint(0, self.n_samples_in_ - 1, (n, self.n_negatives))
                self.neg_indices_ = raw + (raw >= torch.arange(n).unsqueeze(1)).long()
    Ze = cls(**common).fit_transform(X)
    assert torch.equal(Ze, Zr), float((Ze - Zr).abs().max())


def test_eval_metrics_host_flow_equals_reference(ref, monkeypatch):
    """neighborhood_preservation / knn_label_accuracy (torchdr/eval): same values as the reference's CPU evaluation,
    per sample, through the engine's seams on the CPU stand-ins."""
    import fake_ops
    from torchdr.eval import knn_label_accuracy as ref_acc
    from torchdr.eval import neighborhood_preservation as ref_np

    import torchdr_b200 as tb
    from torchdr_b200 import eval as tb_eval

    fake_ops.install(monkeypatch)
    fake_ops.install_entropic(monkeypatch)
    monkeypatch.setattr(tb_eval, "_to_device_tensor", lambda X, device="auto": torch.as_tensor(X))
    n = 240
    X = _data(n, 14, 51)
    Z = torch.randn(n, 2, generator=torch.Generator().manual_seed(2)) + X[:, :2]
    labels = torch.randint(0, 4, (n,), generator=torch.Generator().manual_seed(3))
    for K in (5, 20):
        a = tb.neighborhood_preservation(X, Z, K=K, return_per_sample=True)
        b = ref_np(X, Z, K=K, backend=None, device="cpu", return_per_sample=True)
        assert torch.equal(a, b)
        assert float(tb.neighborhood_preservation(X, Z, K=K)) == float(ref_np(X, Z, K=K, backend=None, device="cpu"))
    for k in (1, 10):
        a = tb.knn_label_accuracy(X, labels, k=k, return_per_sample=True)
        b = ref_acc(X, labels, k=k, backend=None, device="cpu", return_per_sample=True)
        assert torch.equal(a, b)
    assert isinstance(tb.neighborhood_preservation(X.numpy(), Z.numpy(), K=5), float)


def test_distance_seams_equal_reference(ref, monkeypatch):
    """`pairwise_distances` / `pairwise_distances_indexed` with the reference's signature: every argument combination
    the path uses returns what the reference's backend=None call returns (values, index dtype, the (C, None) rule for
    k >= n, error messages) — the engine's dispatch on the CPU stand-ins."""
    import fake_ops
    from torchdr.distance import pairwise_distances as ref_pd
    from torchdr.distance import pairwise_distances_indexed as ref_pdi

    import torchdr_b200 as tb

    fake_ops.install(monkeypatch)
    fake_ops.install_entropic(monkeypatch)
    X, Y = _data(120, 7, 61), _data(90, 7, 62)
    for metric in ("sqeuclidean", "euclidean"):
        assert torch.equal(tb.pairwise_distances(X, metric=metric), ref_pd(X, metric=metric, backend=None))
        assert torch.equal(tb.pairwise_distances(X, Y, metric=metric), ref_pd(X, Y, metric=metric, backend=None))
        assert torch.equal(tb.pairwise_distances(X, metric=metric, exclude_diag=True),
                           ref_pd(X, metric=metric, backend=None, exclude_diag=True))
        for k in (1, 15):
            C, I = tb.pairwise_distances(X, metric=metric, k=k, exclude_diag=True, return_indices=True)
            Cr, Ir = ref_pd(X, metric=metric, backend=None, k=k, exclude_diag=True, return_indices=True)
            assert torch.equal(C, Cr) and torch.equal(I, Ir) and I.dtype == Ir.dtype
            C, I = tb.pairwise_distances(X, Y, metric=metric, k=k, return_indices=True)
            Cr, Ir = ref_pd(X, Y, metric=metric, backend=None, k=k, return_indices=True)
            assert torch.equal(C, Cr) and torch.equal(I, Ir)
        C, I = tb.pairwise_distances(X, metric=metric, k=500, return_indices=True)  # k >= n: full matrix, no indices
        Cr, Ir = ref_pd(X, metric=metric, backend=None, k=500, return_indices=True)
        assert I is None and Ir is None and torch.equal(C, Cr)
    for fn in (tb.pairwise_distances, lambda *a, **k: ref_pd(*a, backend=None, **k)):
        with pytest.raises(ValueError, match="distance is not supported"):
            fn(X, metric="chebyshev")
    Z = torch.randn(120, 2, generator=torch.Generator().manual_seed(1))
    key = torch.randint(0, 120, (120, 9), generator=torch.Generator().manual_seed(2))
    q = torch.randperm(120, generator=torch.Generator().manual_seed(3))[:50]
    for metric in ("sqeuclidean", "euclidean"):
        assert torch.equal(tb.pairwise_distances_indexed(Z, key_indices=key, metric=metric),
                           ref_pdi(Z, key_indices=key, metric=metric, backend=None))
        assert torch.equal(tb.pairwise_distances_indexed(Z, query_indices=q, key_indices=key[:50], metric=metric),
                           ref_pdi(Z, query_indices=q, key_indices=key[:50], metric=metric, backend=None))


def test_affinity_seams_equal_reference(ref, monkeypatch):
    """`UMAPAffinity` / `EntropicAffinity` called like the reference's classes (affinity/base.py:407-561): return
    values, index dtypes and padding, and the fitted attributes the neighbor-embedding driver reads."""
    import fake_ops
    from torchdr.affinity import EntropicAffinity as RefEA
    from torchdr.affinity import UMAPAffinity as RefUA

    import torchdr_b200 as tb

    fake_ops.install(monkeypatch)
    fake_ops.install_entropic(monkeypatch)
    X = _data(210, 11, 71)
    ra = RefUA(n_neighbors=12, max_iter=100, backend=None, device="cpu")
    Vr, Ir = ra(X, return_indices=True)
    a = tb.UMAPAffinity(n_neighbors=12, max_iter=100)
    V, I = a(X, return_indices=True)
    assert torch.equal(V, Vr) and torch.equal(I, Ir) and I.dtype == Ir.dtype
    assert torch.equal(a.rho_.reshape(-1), ra.rho_.reshape(-1)) and torch.equal(a.eps_.reshape(-1), ra.eps_.reshape(-1))
    assert torch.equal(tb.UMAPAffinity(n_neighbors=12)(X, return_indices=False), RefUA(n_neighbors=12, backend=None, device="cpu")(X, return_indices=False))
    rn = RefUA(n_neighbors=12, max_iter=100, backend=None, device="cpu", symmetrize=False)(X, return_indices=True)
    en = tb.UMAPAffinity(n_neighbors=12, max_iter=100, symmetrize=False)(X, return_indices=True)
    assert torch.equal(en[0], rn[0]) and torch.equal(en[1].long(), rn[1].long())
    re_ = RefEA(perplexity=9, max_iter=100, backend=None, device="cpu")
    ee = tb.EntropicAffinity(perplexity=9, max_iter=100)
    for log in (True, False):
        Pr, Jr = re_(X, log=log, return_indices=True)
        P, J = ee(X, log=log, return_indices=True)
        assert torch.equal(P, Pr) and torch.equal(J, Jr) and J.dtype == Jr.dtype
    assert torch.equal(ee.eps_.reshape(-1), re_.eps_.reshape(-1))
    assert ee.log_normalization_.shape == re_.log_normalization_.shape
    assert torch.equal(ee.log_normalization_, re_.log_normalization_)


def test_pca_init_matches_reference_init(ref, monkeypatch):
    """init="pca" (the estimators' default): the engine's covariance + eigh route against the reference's own
    initialisation (full-SVD PCA with svd_flip, then 1e-4 / std rescale, affinity_matcher.py:535-550), captured from a
    live fit.  Same subspace and sign convention; fp32 round-off of two different factorisations apart."""
    import fake_ops
    import torchdr
    from helpers import rel_fro

    import torchdr_b200 as tb

    X = _data(400, 20, 81)
    got = {}

    def capture(base, tag):
        class Cap(base):
            def on_training_step_start(self):
                if int(self.n_iter_) == 0:
                    got[tag] = self.embedding_.detach().clone()
                super().on_training_step_start()

        return Cap

    capture(torchdr.UMAP, "ref")(n_neighbors=10, max_iter=1, init="pca", backend=None, device="cpu", random_state=0,
                                 process_duplicates=False).fit_transform(X)
    fake_ops.install(monkeypatch)
    capture(tb.UMAP, "eng")(n_neighbors=10, max_iter=1, init="pca", random_state=0,
                            process_duplicates=False).fit_transform(X)
    assert rel_fro(got["eng"], got["ref"]) < 1e-4
    torch.testing.assert_close(got["eng"][:, 0].std(), torch.tensor(1e-4), rtol=1e-4, atol=0)
