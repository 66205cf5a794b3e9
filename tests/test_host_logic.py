"""CPU tests of the host-side mirror: partition arithmetic, parameter clamps, bracket scalars,
optimiser/scheduler sequences, and the world_size-2 collectives under gloo."""

import os

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import oracle
from helpers import golden, t
from torchdr_b200.affinity import check_neighbor_param, entropic_bound_scalars
from torchdr_b200.distributed import DistributedContext, all_bounds, chunk_bounds


def test_chunk_bounds_match_reference_rule():
    # torchdr/tests/test_distributed.py:116-129 overwrites rank/world_size by hand
    ctx = DistributedContext(force_enable=True)
    for n, w in [(100, 4), (103, 4), (10, 3), (7, 8), (10_000_000, 8)]:
        prev = 0
        for r in range(w):
            ctx.rank, ctx.world_size = r, w
            s, e = ctx.compute_chunk_bounds(n)
            assert (s, e) == oracle.chunk_bounds(n, r, w) and s == prev
            prev = e
        assert prev == n
    idx = torch.tensor([0, 25, 50, 75, 99])
    assert DistributedContext.get_rank_for_indices(idx, 100, 4).tolist() == [0, 1, 2, 3, 3]
    idx = torch.arange(103)
    owners = DistributedContext.get_rank_for_indices(idx, 103, 4)
    assert owners.tolist() == [oracle.owner_of(i, 103, 4) for i in range(103)]


def test_neighbor_param_clamp():
    assert check_neighbor_param(30, 2000) == 30
    assert check_neighbor_param(5000, 2000) == 1998
    assert check_neighbor_param(1, 2000) == 2
    with pytest.raises(ValueError, match="less than one sample"):
        check_neighbor_param(5, 1)


def test_entropic_bracket_scalars_reproduce_reference_bounds():
    g = golden("entropic_n300_d16_p10")
    C = t(g["C"])
    b_num, b_den, b_lr, b_logp1 = (torch.tensor(v, dtype=torch.float32) for v in entropic_bound_scalars(300, 10))
    dN, d1, d2 = C.max(1)[0], C[:, 0], C[:, 1]
    beta_lo = torch.max(b_num / (b_den * (dN - d1)), torch.sqrt(b_lr / (dN * dN - d1 * d1)))
    beta_hi = b_logp1 / (d2 - d1)
    assert torch.equal(1 / beta_hi, t(g["begin"])) and torch.equal(1 / beta_lo, t(g["end"]))


def _make(cls, **kw):
    m = cls(**kw)
    m.n_samples_in_ = 300
    m.early_exaggeration_coeff_ = m.early_exaggeration_coeff
    m._dummy = torch.nn.Parameter(torch.zeros(1))
    m.params_ = [{"params": [m._dummy]}]
    m._set_learning_rate()
    m._configure_optimizer()
    m._configure_scheduler()
    return m


def test_schedules_match_reference_runs():
    import torchdr_b200 as tb

    g = golden("umap_n300_d16_k15")
    m = _make(tb.UMAP, n_neighbors=15, max_iter=100)
    lrs = []
    for _ in range(100):
        lrs.append(m._hyper()[0])
        m._advance_schedule()
    np.testing.assert_array_equal(np.asarray(lrs, dtype=np.float32), g["lr"].astype(np.float32))
    assert m._hyper()[1] == 0.0  # plain SGD (umap.py:139)

    g = golden("largevis_n300_d16_p10")
    m = _make(tb.LargeVis, perplexity=10, max_iter=30)
    lrs = []
    for _ in range(30):
        lrs.append(m._hyper()[0])
        m._advance_schedule()
    np.testing.assert_allclose(lrs, g["lr"], rtol=1e-12)
    assert m._hyper()[1] == 0.8

    # TSNE: after the early-exaggeration rebuild lr and momentum keep their first-build values
    m = _make(tb.TSNE, perplexity=10, max_iter=20, early_exaggeration_iter=10)
    assert m._hyper() == (50.0, 0.5)
    m.early_exaggeration_coeff_ = 1
    m._set_learning_rate()
    m._configure_optimizer()
    m._configure_scheduler()
    assert m.lr_ == 75.0 and m._hyper() == (50.0, 0.5)  # verified on the reference (oracle/tsne.py)


def test_constructor_surface_and_errors():
    import torchdr_b200 as tb

    m = tb.UMAP()
    assert (m.n_neighbors, m.max_iter, m.lr, m.n_negatives, m.init) == (30, 1000, 1.0, 150, "pca")
    assert abs(m._a - 1.5769434602697652) < 1e-12 and abs(m._b - 0.8950608778515733) < 1e-12
    assert tb.TSNE().early_exaggeration_coeff == 12.0 and tb.LargeVis().n_negatives == 5
    it, sn = tb.InfoTSNE(), tb.SNE()  # infotsne.py:106-136, sne.py:95-120
    assert (it.n_negatives, it.early_exaggeration_coeff, it.early_exaggeration_iter, it.max_iter, it.scheduler) == \
        (300, 12, 250, 1000, "LinearLR")
    assert (sn.early_exaggeration_coeff, sn.early_exaggeration_iter, sn.max_iter, sn.scheduler, sn.lr) == \
        (1, 0, 2000, None, "auto")
    with pytest.raises(ValueError, match="distance is not supported"):
        tb.UMAPAffinity(metric="chebyshev")
    with pytest.raises(RuntimeError, match="requires launching with torchrun"):
        tb.UMAP(distributed=True)
    with pytest.raises(ValueError, match="not fitted yet"):
        tb.UMAP().transform()


# ----------------------------------------------------------------------------- gloo, world_size 2
def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from torchdr_b200.distributed import all_gather_rows, exchange_edges

        # edge exchange: rank r sends (row, col, val) triples to the owner of row
        n = 11
        bounds = all_bounds(n, world)
        g = torch.Generator().manual_seed(rank)
        rows = torch.randint(0, n, (20,), generator=g)
        owners = DistributedContext.get_rank_for_indices(rows, n, world)
        keep = owners != rank
        rows, owners = rows[keep], owners[keep]
        order = torch.argsort(owners, stable=True)
        rows = rows[order]
        counts = torch.bincount(owners, minlength=world)
        cols = (rows * 7 + rank).int()
        vals = rows.float() + 0.5 * rank
        rr, rc, rv = exchange_edges(counts, rows, cols, vals)
        s, e = bounds[rank]
        ok = bool(((rr >= s) & (rr < e)).all()) and rr.dtype == torch.int64 and rc.dtype == torch.int32
        ok = ok and bool((rc == (rr * 7 + (1 - rank)).int()).all()) and bool((rv == rr.float() + 0.5 * (1 - rank)).all())
        # embedding exchange: uneven chunks (6 + 5 rows)
        Z = torch.full((n, 2), -1.0)
        Z[s:e] = torch.arange(s, e, dtype=torch.float32)[:, None] + torch.tensor([0.0, 0.25])
        all_gather_rows(Z, bounds, rank)
        want = torch.arange(n, dtype=torch.float32)[:, None] + torch.tensor([0.0, 0.25])
        ok = ok and torch.equal(Z, want)
        # equal chunks take the in-place single-collective path
        b10 = all_bounds(10, world)
        s10, e10 = b10[rank]
        Z10 = torch.full((10, 2), -1.0)
        Z10[s10:e10] = torch.arange(s10, e10, dtype=torch.float32)[:, None] + torch.tensor([0.0, 0.5])
        all_gather_rows(Z10, b10, rank)
        ok = ok and torch.equal(Z10, torch.arange(10, dtype=torch.float32)[:, None] + torch.tensor([0.0, 0.5]))
        # index >= 2^24 survives the exchange (the reference casts indices to fp32, sparse.py:286-293)
        big = torch.tensor([2**24 + 1 + rank], dtype=torch.int64)
        cnt = torch.zeros(world, dtype=torch.int64)
        cnt[1 - rank] = 1
        br, _, _ = exchange_edges(cnt, big, big.int(), big.float())
        ok = ok and int(br[0]) == 2**24 + 1 + (1 - rank)
        q.put((rank, ok))
    finally:
        dist.destroy_process_group()


def test_collectives_world_size_2_gloo():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
    assert res == [(0, True), (1, True)]


def test_tile_prune_rule_is_exact_on_cpu():
    """The rule the CUDA pruned sweep implements (oracle/knn.py:tile_prune_plan, same margins): kNN restricted to
    the surviving tiles equals the dense kNN of the reference path, on index-local clusters (where it prunes),
    on shuffled rows (where it cannot) and with clusters smaller than a tile (every tile straddles several)."""
    import oracle
    from helpers import clustered

    for n, d, k, shuffle in ((6000, 32, 15, False), (6000, 32, 15, True), (4000, 16, 5, False)):
        X = clustered(n, d)  # min(1000, n // 100) clusters of ~100 points, contiguous
        if shuffle:
            X = X[torch.randperm(n, generator=torch.Generator().manual_seed(0))].contiguous()
        keep, tau = oracle.tile_prune_plan(X, k)
        C_ref, I_ref = oracle.knn_dense(X, k)
        C, I = oracle.knn_with_tile_mask(X, k, keep)
        idx64, _, entry_ok, _ = oracle.knn_ambiguity(X, k)
        assert torch.equal(I.long()[entry_ok], I_ref.long()[entry_ok])
        # blockwise and full sgemm may round differently: distances on the scale of the norms (knn_ambiguity's gap)
        assert float((C - C_ref).abs().max()) <= 4e-6 * 2 * float((X**2).sum(1).max())
        assert bool((tau >= C_ref[:, k - 1] - 1e-3).all())  # tau bounds the k-th neighbour distance
        frac = float(keep.float().mean())
        assert bool(keep.diagonal().all())
        if not shuffle:
            assert frac < 0.5, frac
        else:
            assert frac > 0.9, frac


def test_knn_options_validate_without_a_gpu():
    """path / prune are per-call arguments of the C ABI (no process-wide switches): bad values are rejected before any
    device work, so this runs without a GPU."""
    import ctypes

    from torchdr_b200 import _lib

    lib = _lib.load()
    one = ctypes.c_void_p(256)  # never dereferenced: argument validation comes first
    args = (one, 4, 0, one, 4, 8, 2, 1, 0, one, one)
    assert lib.tdr_knn_f32(*args, 7, -1, None, None, None, 0, None) == _lib.TDR_E_INVALID
    assert "path" in _lib.last_error()
    assert lib.tdr_knn_f32(*args, 0, 3, None, None, None, 0, None) == _lib.TDR_E_INVALID
    assert "prune" in _lib.last_error()
    assert not hasattr(lib, "tdr_knn_set_prune_") and "tdr_knn_set_prune" not in _lib.SIGNATURES
    # the workspace query covers the pruned sweep's buffers (boxes, bounds, tile lists) once there are >= 64 tiles
    small = lib.tdr_knn_workspace_bytes(64 * 128 - 1, 64 * 128 - 128, 128, 15)
    big = lib.tdr_knn_workspace_bytes(64 * 128, 64 * 128, 128, 15)
    assert big > small + 64 * 64 * 4


def test_voronoi_tree_order_restores_locality_on_cpu():
    """torchdr_b200/reorder.py (experimental): a valid permutation; on shuffled clustered rows the re-ordered tiles
    are spatially compact again, so the pruning rule (oracle/knn.py:tile_prune_plan) drops most tile pairs, and a kNN
    run on the re-ordered rows maps back to the kNN of the original rows."""
    import oracle
    from helpers import clustered
    from torchdr_b200.reorder import knn_in_any_order, voronoi_tree_order

    n, d, k = 6000, 32, 10
    X = clustered(n, d)
    Xs = X[torch.randperm(n, generator=torch.Generator().manual_seed(0))].contiguous()
    g = torch.Generator().manual_seed(1)
    perm = voronoi_tree_order(Xs, generator=g)
    assert perm.dtype == torch.long and torch.equal(perm.sort().values, torch.arange(n))
    keep_shuffled, _ = oracle.tile_prune_plan(Xs, k)
    keep_sorted, _ = oracle.tile_prune_plan(Xs[perm], k)
    assert float(keep_shuffled.float().mean()) > 0.9
    assert float(keep_sorted.float().mean()) < 0.35, float(keep_sorted.float().mean())
    # duplicates cannot be split: the recursion must still terminate and return a permutation
    Xdup = torch.cat([Xs[:300], Xs[:1].repeat(700, 1)])
    pd = voronoi_tree_order(Xdup, generator=g)
    assert torch.equal(pd.sort().values, torch.arange(1000))
    # kNN through the permutation == kNN of the original rows
    C_ref, I_ref = oracle.knn_dense(Xs, k)
    C, I, _ = knn_in_any_order(Xs, k, lambda Xp: oracle.knn_dense(Xp, k), perm=perm)
    _, _, entry_ok, _ = oracle.knn_ambiguity(Xs, k)
    assert torch.equal(I.long()[entry_ok], I_ref.long()[entry_ok])
    assert float((C - C_ref).abs().max()) <= 4e-6 * 2 * float((Xs**2).sum(1).max())


def test_umap_loop_bookkeeping_without_a_gpu(monkeypatch):
    """The batched UMAP loop (torchdr_b200/neighbor_embedding.py:UMAP._loop) with the native call replaced by a
    recorder: batches end at the reference's check points (affinity_matcher.py:331-349), the learning rates handed
    to the kernels are the reference's LinearLR sequence although they are produced one batch ahead, and the loop
    stops at the first check whose gradient norm is below min_grad_norm."""
    import oracle
    from torchdr_b200 import neighbor_embedding as ne
    from torchdr_b200 import ops

    calls = []

    def fake_run(Za, Zb, rowptr, col, eps, eons, n_iter0, lrs, a, b, gnorm_sq=None, **kw):
        calls.append((int(n_iter0), [float(x) for x in lrs], gnorm_sq is not None))
        if gnorm_sq is not None:
            gnorm_sq.fill_(fake_run.gnorm_sq)
        return Za if len(lrs) % 2 == 0 else Zb

    monkeypatch.setattr(ops, "umap_run", fake_run)

    def run(max_iter, check_interval, gnorm_sq, min_grad_norm=1e-7):
        calls.clear()
        fake_run.gnorm_sq = gnorm_sq
        m = ne.UMAP(n_neighbors=5, max_iter=max_iter, check_interval=check_interval, min_grad_norm=min_grad_norm,
                    a=1.5, b=0.9, random_state=0, distributed=False)
        m.n_samples_in_, m.chunk_start_, m.chunk_end_ = 8, 0, 8
        m.early_exaggeration_coeff_ = 1
        m._native_opt = m._uses_native_sgd()
        m._graph = (torch.zeros(9, dtype=torch.long), torch.zeros(0, dtype=torch.int32), torch.zeros(0), torch.zeros(0))
        m.embedding_ = torch.zeros(8, 2)
        m._gnorm, m._nan = torch.zeros(1, dtype=torch.float64), torch.zeros(1, dtype=torch.int32)
        m._dummy = torch.nn.Parameter(torch.zeros(1))
        m.params_ = [{"params": [m._dummy]}]
        m._set_learning_rate()
        m._configure_optimizer()
        m._configure_scheduler()
        m._loop()
        return m

    m = run(130, 50, gnorm_sq=1.0)
    assert [(c[0], len(c[1]), c[2]) for c in calls] == [(0, 1, True), (1, 50, True), (51, 50, True), (101, 29, False)]
    lrs = np.asarray([x for c in calls for x in c[1]], dtype=np.float32)
    assert np.array_equal(lrs, oracle.linear_lr_sequence(1.0, 130, 130))  # umap.py:139, NE base.py:175-182
    assert m._last_step == 129
    m = run(400, 20, gnorm_sq=1e-20)  # converged at the very first check
    assert [(c[0], len(c[1])) for c in calls] == [(0, 1)] and m._last_step == 0
    m = run(75, 25, gnorm_sq=1.0)
    assert [(c[0], len(c[1])) for c in calls] == [(0, 1), (1, 25), (26, 25), (51, 24)]


def test_umap_early_exaggeration_switch_without_a_gpu(monkeypatch):
    """UMAP(early_exaggeration_coeff > 1, early_exaggeration_iter = k), accepted through **kwargs as in the reference:
    iterations 0..k run with lambda = coeff, then the coefficient drops to 1 and optimiser + scheduler are rebuilt
    (NE base.py:282-295) — the batch is split at k and the learning-rate sequence restarts from the new scheduler."""
    from torchdr_b200 import neighbor_embedding as ne
    from torchdr_b200 import ops

    calls = []

    def fake_run(Za, Zb, rowptr, col, eps, eons, n_iter0, lrs, a, b, gnorm_sq=None, lam=1.0, **kw):
        calls.append((int(n_iter0), len(lrs), float(lam), [float(x) for x in lrs]))
        if gnorm_sq is not None:
            gnorm_sq.fill_(1.0)
        return Za if len(lrs) % 2 == 0 else Zb

    monkeypatch.setattr(ops, "umap_run", fake_run)
    m = ne.UMAP(n_neighbors=5, max_iter=40, check_interval=10, a=1.5, b=0.9, random_state=0, distributed=False,
                early_exaggeration_coeff=4.0, early_exaggeration_iter=13)
    m.n_samples_in_, m.chunk_start_, m.chunk_end_ = 8, 0, 8
    m.early_exaggeration_coeff_ = m.early_exaggeration_coeff
    m._native_opt = m._uses_native_sgd()
    m._graph = (torch.zeros(9, dtype=torch.long), torch.zeros(0, dtype=torch.int32), torch.zeros(0), torch.zeros(0))
    m.embedding_ = torch.zeros(8, 2)
    m._gnorm, m._nan = torch.zeros(1, dtype=torch.float64), torch.zeros(1, dtype=torch.int32)
    m._dummy = torch.nn.Parameter(torch.zeros(1))
    m.params_ = [{"params": [m._dummy]}]
    m._set_learning_rate()
    m._configure_optimizer()
    m._configure_scheduler()
    m._loop()
    assert [(c[0], c[1], c[2]) for c in calls] == [(0, 1, 4.0), (1, 10, 4.0), (11, 3, 4.0), (14, 7, 1.0), (21, 10, 1.0),
                                                   (31, 9, 1.0)]
    lrs = [x for c in calls for x in c[3]]
    # the reference's own objects, rebuilt the reference's way (ONE params_ dict reused by the rebuild,
    # affinity_matcher.py:588-590: torch keeps the group's current lr / initial_lr, so the new LinearLR continues from
    # the lr reached at the switch instead of restarting at 1 — the quirk DESIGN.md section 1 documents)
    dummy = torch.nn.Parameter(torch.zeros(1))
    params = [{"params": [dummy]}]
    kw = {"start_factor": torch.tensor(1.0), "end_factor": torch.tensor(0), "total_iters": 40}
    opt = torch.optim.SGD(params, lr=1.0)
    sch = torch.optim.lr_scheduler.LinearLR(opt, **kw)
    expect = []
    for t in range(40):
        expect.append(float(opt.param_groups[0]["lr"]))
        opt.step()
        sch.step()
        if t == 13:
            opt = torch.optim.SGD(params, lr=1.0)
            sch = torch.optim.lr_scheduler.LinearLR(opt, **kw)
    np.testing.assert_allclose(lrs, expect, rtol=1e-6)
    assert abs(lrs[14] - lrs[13] * (1 - 1 / 40) / 1.0) < 0.05 and lrs[14] < 0.7  # continues, does not restart at 1
    assert m.early_exaggeration_coeff_ == 1 and m._last_step == 39


def test_momentum_loop_bookkeeping_matches_reference_sequences(monkeypatch):
    """The gradient + momentum-SGD loop shared by LargeVis / TSNE / InfoTSNE / SNE with the two native calls replaced
    by recorders: the (lr, momentum, buffer-restart) triple handed to `tdr_sgd_momentum_f32` at every step must be the
    sequence the REFERENCE run produced (lr captured in tests/golden/*.npz by make_golden*.py), including the quirk
    that the optimiser rebuilt after early exaggeration keeps the first build's lr and momentum
    (affinity_matcher.py:588-590, oracle/tsne.py) and the LinearLR restart of InfoTSNE."""
    from torchdr_b200 import neighbor_embedding as ne
    from torchdr_b200 import ops

    steps = []
    monkeypatch.setattr(ops, "sgd_momentum",
                        lambda Z, mom, grad, lr, mu, first, **kw: steps.append((float(lr), float(mu), bool(first))))

    def run(cls, name, **kw):
        steps.clear()
        g = golden(name)
        m = cls(perplexity=10, init="normal", random_state=0, min_grad_norm=0.0, distributed=False, **kw)
        m._compute_gradient = lambda Z, step: None
        m.n_samples_in_, m.chunk_start_, m.chunk_end_ = 300, 0, 300
        m.early_exaggeration_coeff_ = m.early_exaggeration_coeff
        m._native_opt = m._uses_native_sgd()
        m.embedding_ = torch.zeros(300, 2)
        m._gnorm, m._nan = torch.ones(1, dtype=torch.float64), torch.zeros(1, dtype=torch.int32)
        m._dummy = torch.nn.Parameter(torch.zeros(1))
        m.params_ = [{"params": [m._dummy]}]
        m._set_learning_rate()
        m._configure_optimizer()
        m._configure_scheduler()
        m._loop()
        lr_ref = np.asarray(g["lr"], dtype=np.float64)
        assert len(steps) == len(lr_ref)
        np.testing.assert_allclose([s[0] for s in steps], lr_ref, rtol=1e-7)
        return [s[1] for s in steps], [s[2] for s in steps]

    mu, first = run(ne.LargeVis, "largevis_n300_d16_p10", max_iter=30)
    assert set(mu) == {0.8} and first == [True] + [False] * 29
    mu, first = run(ne.TSNE, "tsne_n300_d16_p10", max_iter=20, early_exaggeration_iter=10)
    assert set(mu) == {0.5}  # the rebuilt optimiser keeps momentum 0.5 (and lr 50)
    assert [i for i, f in enumerate(first) if f] == [0, 11]  # momentum buffer restarts after the switch at step 10
    mu, first = run(ne.InfoTSNE, "infotsne_n300_d16_p10", max_iter=20, early_exaggeration_iter=10, n_negatives=50)
    assert set(mu) == {0.5} and [i for i, f in enumerate(first) if f] == [0, 11]
    mu, first = run(ne.SNE, "sne_n300_d16_p10", max_iter=20)
    assert set(mu) == {0.8} and first == [True] + [False] * 19


def test_umap_estimator_host_flow_reproduces_reference_run_on_cpu(monkeypatch):
    """The public estimator with the native calls replaced by oracle-backed stand-ins (tests/fake_ops.py): what is
    left under test is the HOST code — fit_transform, UMAPAffinity, the edge schedule hand-over, the per-step hook
    path and the batched path, the LinearLR bookkeeping.  Driven like the reference's golden run (same init, same
    negatives through the hook) it must land on the reference's embedding bit for bit."""
    import fake_ops
    from helpers import negative_table

    import torchdr_b200 as tb

    fake_ops.install(monkeypatch)
    g = golden("umap_n300_d16_k15")
    seed = int(g["seed"])
    snaps = {}

    class Injected(tb.UMAP):
        def on_training_step_start(self):
            self.neg_indices_ = negative_table(seed, int(self.n_iter_), 300, 75)

        def on_training_step_end(self):
            if int(self.n_iter_) + 1 in (1, 2, 5, 20):
                snaps[int(self.n_iter_) + 1] = self.embedding_.clone()

    m = Injected(n_neighbors=15, max_iter=int(g["max_iter"]), init=t(g["Zinit"]), random_state=0,
                 process_duplicates=False, min_grad_norm=0.0, check_interval=20)
    # stop after the 20th step without touching max_iter (it sets the schedule and the edge threshold)
    m._converged = lambda step, gn: step >= 20
    Z = m.fit_transform(t(g["X"]).numpy())
    assert isinstance(Z, np.ndarray) and Z.shape == (300, 2)
    for T in (1, 2, 5, 20):
        assert torch.equal(snaps[T], t(g[f"Z_{T}"])), T
    # batched path (no hooks): in-kernel negatives are the stand-in's own draws, so only the bookkeeping is checked —
    # finite result, iteration count, early stop at the first check
    m2 = tb.UMAP(n_neighbors=15, max_iter=60, init="normal", random_state=0, process_duplicates=False, check_interval=25)
    Z2 = m2.fit_transform(t(g["X"]))
    assert Z2.shape == (300, 2) and bool(torch.isfinite(Z2).all()) and int(m2.n_iter_) == 59
    m3 = tb.UMAP(n_neighbors=15, max_iter=60, init="normal", random_state=0, process_duplicates=False,
                 check_interval=25, min_grad_norm=1e9)
    m3.fit_transform(t(g["X"]))
    assert int(m3.n_iter_) == 0


def test_momentum_estimators_host_flow_reproduce_reference_runs_on_cpu(monkeypatch):
    """Same arrangement for the autograd-mode estimators: public API + host loop (early exaggeration switch, optimiser
    rebuild, LinearLR, momentum buffer) on oracle-backed stand-ins must reproduce the reference's LargeVis, t-SNE,
    InfoTSNE and SNE runs captured in tests/golden bit for bit (same init; negatives injected through the hook)."""
    import fake_ops
    from helpers import negative_table

    import torchdr_b200 as tb

    fake_ops.install(monkeypatch)
    fake_ops.install_entropic(monkeypatch)

    def snapshots(cls, g, steps, **kw):
        snaps = {}
        n_neg = int(g["n_neg"]) if "n_neg" in g else 5

        class Captured(cls):
            def on_training_step_start(self):
                if "seed" in g:
                    self.neg_indices_ = negative_table(int(g["seed"]), int(self.n_iter_), 300, n_neg)

            def on_training_step_end(self):
                if int(self.n_iter_) + 1 in steps:
                    snaps[int(self.n_iter_) + 1] = self.embedding_.clone()

        m = Captured(perplexity=10, init=t(g["Zinit"]), random_state=0, process_duplicates=False, min_grad_norm=0.0, **kw)
        m.fit_transform(t(g["X"]))
        return snaps

    g = golden("largevis_n300_d16_p10")
    snaps = snapshots(tb.LargeVis, g, (1, 2, 5, 10, 30), max_iter=30)
    for T in (1, 2, 5, 10, 30):
        assert torch.equal(snaps[T], t(g[f"Z_{T}"])), ("largevis", T)
    g = golden("tsne_n300_d16_p10")
    snaps = snapshots(tb.TSNE, g, (1, 2, 5, 10, 11, 12, 20), max_iter=20, early_exaggeration_iter=10)
    for T in (1, 2, 5, 10, 11, 12, 20):
        assert torch.equal(snaps[T], t(g[f"Z_{T}"])), ("tsne", T)
    g = golden("infotsne_n300_d16_p10")
    snaps = snapshots(tb.InfoTSNE, g, (1, 2, 5, 10, 11, 12, 20), max_iter=20, early_exaggeration_iter=10, n_negatives=50)
    for T in (1, 2, 5, 10, 11, 12, 20):
        assert torch.equal(snaps[T], t(g[f"Z_{T}"])), ("infotsne", T)
    g = golden("sne_n300_d16_p10")
    snaps = snapshots(tb.SNE, g, (1, 2, 5, 10, 20), max_iter=20)
    for T in (1, 2, 5, 10, 20):
        assert torch.equal(snaps[T], t(g[f"Z_{T}"])), ("sne", T)


def test_reordered_affinity_equals_plain_affinity_on_cpu(monkeypatch):
    """knn_order="tree": searching in the Voronoi-tree order and mapping the rows back must give the same symmetrised
    graph as searching in the input order (host logic on the CPU stand-ins)."""
    import fake_ops

    import torchdr_b200 as tb
    from torchdr_b200 import ops

    from torchdr_b200 import reorder

    fake_ops.install(monkeypatch)
    monkeypatch.setattr(reorder, "MIN_ROWS_FOR_REORDER", 0)
    g = golden("umap_n300_d16_k15")
    X = t(g["X"])
    plain = tb.UMAPAffinity(n_neighbors=15, max_iter=100, knn_order="input").compute_csr(X)
    aff = tb.UMAPAffinity(n_neighbors=15, max_iter=100, knn_order="tree")
    reordered = aff.compute_csr(X)
    for a, b in zip(plain, reordered):
        assert torch.equal(a, b)
    assert torch.equal(aff.eps_, t(g["sigma"])) and torch.equal(aff.rho_, t(g["rho"]))
    vals, idx = aff(X)
    assert torch.equal(idx, t(g["sym_idx"]).long()) and torch.equal(vals, t(g["sym_vals"]))


def test_fit_in_tree_order_returns_rows_in_input_order_on_cpu(monkeypatch):
    """Estimator-level re-ordering (neighbor_embedding._fit_order): a fit on rows WITHOUT index locality runs on
    X[perm] and must hand back row r = the embedding of input row r.  With an injected initialisation and zero
    optimisation steps' worth of randomness removed (max_iter=1, negatives from the stand-in's own seeded stream are
    index-dependent, so the check is on the graph-independent part): the initial layout is un-permuted exactly, and the
    locality probe takes its three decisions (clustered order kept, shuffled re-ordered, structureless kept)."""
    import fake_ops

    import torchdr_b200 as tb
    from torchdr_b200 import reorder

    g = torch.Generator().manual_seed(0)
    nc, per, d = 24, 256, 16
    centers = torch.randn(nc, d, generator=g) * 10
    Xc = (centers.repeat_interleave(per, 0) + 0.5 * torch.randn(nc * per, d, generator=g)).contiguous()
    shuf = torch.randperm(nc * per, generator=g)
    Xs = Xc[shuf].contiguous()
    Xu = torch.randn(nc * per, d, generator=g)
    assert reorder.index_locality(Xc) < 0.3 < reorder.LOCALITY_THRESHOLD < reorder.index_locality(Xs)
    monkeypatch.setattr(reorder, "MIN_ROWS_FOR_REORDER", 0)
    assert reorder.choose_order(Xc) is None                      # locality already there
    perm = reorder.choose_order(Xs)
    assert perm is not None and torch.equal(torch.sort(perm).values, torch.arange(nc * per))
    assert reorder.index_locality(Xs, perm=perm) < 0.5 * reorder.index_locality(Xs)
    assert reorder.choose_order(Xu) is None                      # no cluster structure: no order helps

    fake_ops.install(monkeypatch)
    seen = {}
    orig = tb.UMAPAffinity.compute_csr

    def spy(self, X):
        seen["X"], seen["order"] = X.clone(), self.knn_order
        return orig(self, X)

    monkeypatch.setattr(tb.UMAPAffinity, "compute_csr", spy)
    Z0 = torch.randn(nc * per, 2, generator=g)
    m = tb.UMAP(n_neighbors=10, max_iter=1, lr=0.0, init=Z0, init_scaling=1.0, random_state=0, process_duplicates=False,
                distributed=False)
    Z = m.fit_transform(Xs)
    assert seen["order"] == "presorted" and torch.equal(seen["X"], Xs[perm])
    # lr = 0: the step leaves the initialisation untouched, so the output must be Z0 (rescaled) in INPUT order
    torch.testing.assert_close(Z, Z0 / Z0[perm][:, 0].std(), rtol=1e-6, atol=0)
    m2 = tb.UMAP(n_neighbors=10, max_iter=1, lr=0.0, init=Z0, init_scaling=1.0, random_state=0, process_duplicates=False,
                 distributed=False, knn_order="input")
    m2.fit_transform(Xs)
    assert seen["order"] == "input" and torch.equal(seen["X"], Xs)


def test_estimators_are_sklearn_estimators_and_torch_modules():
    """base.py:27 ``DRModule(BaseEstimator, nn.Module, ABC)``: get_params / set_params / clone round trips and the module
    tree (the input affinity is a sub-module), for every estimator of the path.  UMAP goes beyond the reference here:
    torchdr.UMAP stores a / b only as _a / _b, so its get_params() raises AttributeError."""
    import torch.nn as nn
    from sklearn.base import BaseEstimator, clone

    import torchdr_b200 as tb

    for cls, kw in ((tb.UMAP, dict(n_neighbors=12, min_dist=0.2)), (tb.TSNE, dict(perplexity=20)),
                    (tb.LargeVis, dict(perplexity=25, n_negatives=7)), (tb.InfoTSNE, dict(perplexity=10)),
                    (tb.SNE, dict(perplexity=10))):
        m = cls(max_iter=17, random_state=3, distributed=False, **kw)
        assert isinstance(m, BaseEstimator) and isinstance(m, nn.Module)
        p = m.get_params()
        assert p["max_iter"] == 17 and p["random_state"] == 3
        for key, val in kw.items():
            assert p[key] == val
        c = clone(m)
        assert type(c) is cls and c is not m and c.get_params().keys() == p.keys()
        for key in p:
            a, b = p[key], c.get_params()[key]
            assert (a is b) or (a == b) or (isinstance(a, dict) and a.keys() == b.keys()), key
        m.set_params(max_iter=5)
        assert m.max_iter == 5
        assert [name for name, _ in m.named_modules()] == ["", "affinity_in"]
        assert isinstance(m.affinity_in, nn.Module)
        assert cls.__name__ in repr(m)
    with pytest.raises(ValueError):
        tb.UMAP(distributed=False).set_params(no_such_parameter=1)
    # dense optimisation is outside the path: refused with a clear message instead of an AttributeError deep inside
    with pytest.raises(NotImplementedError, match="sparsity=False"):
        tb.TSNE(sparsity=False, distributed=False)._compute_affinity(torch.zeros(8, 3))


def test_largevis_row_local_form_equals_scatter_form_on_cpu(monkeypatch):
    """The row-local LargeVis step (union graph S = P + P^T gathered per row + both halves of every negative pair,
    momentum SGD on the local rows: tdr_largevis_step_f32) must be the same optimisation as the reference's
    formulation (autograd scatter of largevis.py:181-201 + SGD on all rows).  Both run here on the CPU stand-ins with
    the same negative stream; the stand-in of the scatter form differentiates the oracle's loss with autograd."""
    import fake_ops
    from helpers import rel_fro

    import torchdr_b200 as tb

    fake_ops.install(monkeypatch)
    fake_ops.install_entropic(monkeypatch)
    g = golden("largevis_n300_d16_p10")
    X, Z0 = t(g["X"]), t(g["Z0"])
    out = {}
    for row_local in (True, False):
        m = tb.LargeVis(perplexity=10, max_iter=12, init=Z0, init_scaling=float(Z0[:, 0].std()), random_state=0,
                        process_duplicates=False, distributed=False, row_local=row_local, knn_order="input")
        out[row_local] = m.fit_transform(X)
        assert (m._row_local() is row_local)
    assert rel_fro(out[True], out[False]) < 1e-5, rel_fro(out[True], out[False])


def test_baseline_config_1_host_flow_on_cpu(monkeypatch):
    """BASELINE.json configs[0] through the public estimator on the CPU stand-ins: TSNE(perplexity=30) on the
    reference's 2000 x 50 blobs run.  Tolerances as in tests/test_oracle_golden.py (the reference's own backward is not
    bit-reproducible at this size)."""
    import fake_ops
    from helpers import rel_fro

    import torchdr_b200 as tb

    fake_ops.install(monkeypatch)
    fake_ops.install_entropic(monkeypatch)
    g = golden("tsne_c1_n2000_d50_p30")
    snaps = {}

    class Captured(tb.TSNE):
        def on_training_step_end(self):
            if int(self.n_iter_) + 1 in (1, 5, 20):
                snaps[int(self.n_iter_) + 1] = self.embedding_.clone()

    m = Captured(perplexity=30, max_iter=20, init=t(g["Zinit"]), random_state=0, process_duplicates=False,
                 min_grad_norm=0.0)
    Z = m.fit_transform(g["X"])
    assert isinstance(Z, np.ndarray) and Z.shape == (2000, 2)
    for T, tol in ((1, 2e-6), (5, 5e-6), (20, 2e-5)):
        assert rel_fro(snaps[T], g[f"Z_{T}"]) < tol, T


def test_duplicate_precheck_matches_torch_unique():
    """The hash pre-check that spares the reference's torch.unique(X, dim=0) (base.py:132-146) when all rows differ:
    it may only say "no duplicates" when unique() would keep every row."""
    from torchdr_b200.neighbor_embedding import _may_have_duplicate_rows

    g = torch.Generator().manual_seed(0)
    for n, d in ((1, 3), (2, 1), (1000, 7), (5000, 128)):
        X = torch.randn(n, d, generator=g)
        assert not _may_have_duplicate_rows(X, chunk=1024)
        assert torch.unique(X, dim=0).shape[0] == n
    X = torch.randn(3000, 16, generator=g)
    Xd = X.clone()
    Xd[2999] = Xd[17]
    assert _may_have_duplicate_rows(Xd, chunk=1000)
    Xz = X.clone()
    Xz[5, 3], Xz[6] = 0.0, Xz[5]
    Xz[6, 3] = -0.0  # unique() compares values: -0.0 == +0.0
    assert torch.unique(Xz, dim=0).shape[0] == 2999 and _may_have_duplicate_rows(Xz)
    Xc = torch.zeros(50, 4)  # all equal
    assert _may_have_duplicate_rows(Xc)
    assert not _may_have_duplicate_rows(torch.arange(40.0).reshape(40, 1))


def test_duplicates_share_their_embedding_on_cpu(monkeypatch):
    """process_duplicates=True (the default) through the public estimator on the CPU stand-ins: duplicates are found,
    the fit runs on the unique rows in unique()'s order and duplicates share their embedding (base.py:132-146); without
    duplicates the rows are fitted in their own order."""
    import fake_ops

    import torchdr_b200 as tb

    fake_ops.install(monkeypatch)
    g = golden("umap_n300_d16_k15")
    X = t(g["X"])
    Xdup = torch.cat([X, X[:25]])
    Z = tb.UMAP(n_neighbors=15, max_iter=12, init="normal", random_state=0, check_interval=5).fit_transform(Xdup)
    assert Z.shape == (325, 2) and torch.equal(Z[:25], Z[300:])
    calls = []
    orig = torch.unique
    monkeypatch.setattr(torch, "unique", lambda *a, **k: (calls.append(1) if k.get("dim") == 0 else None) or orig(*a, **k))
    tb.UMAP(n_neighbors=15, max_iter=3, init="normal", random_state=0).fit_transform(X)
    assert not calls  # no duplicates: the lexicographic row sort is never run


def _sharded_umap_worker(rank, world, port, q):
    """One rank of a world_size-2 gloo run of the PUBLIC estimator on the CPU stand-ins."""
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import fake_ops
        from helpers import negative_table

        import torchdr_b200 as tb

        mp_ = pytest.MonkeyPatch()
        fake_ops.install(mp_)
        fake_ops.install_sharded(mp_)
        g = golden("umap_n300_d16_k15")
        seed = int(g["seed"])
        n = 299  # uneven chunks (150 + 149) on purpose
        X = t(g["X"])[:n].contiguous()

        class Injected(tb.UMAP):
            def on_training_step_start(self):
                s, e = self.chunk_start_, self.chunk_end_
                self.neg_indices_ = negative_table(seed, int(self.n_iter_), n, 75)[s:e].contiguous()

        opt = {} if port % 2 == 0 else {"optimizer": "Adam", "lr": 0.05}  # second scenario: all-reduced gradient path
        m = Injected(n_neighbors=15, max_iter=4, init=t(g["Zinit"])[:n], random_state=0, process_duplicates=False,
                     min_grad_norm=0.0, check_interval=2, **opt)
        assert m.distributed and m.world_size == world and m.rank == rank
        Z = m.fit_transform(X)
        q.put((rank, Z.numpy().tobytes()))
        mp_.undo()
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("scenario", [0, 1], ids=["fused-sgd-allgather", "adam-allreduce"])
def test_sharded_umap_estimator_equals_single_process_under_gloo(scenario, monkeypatch):
    """Two gloo ranks run `UMAP(...).fit_transform` (distributed="auto") on the CPU stand-ins, rows sharded 150 + 149:
    chunked affinity, edge exchange (all-to-all), A_max all-reduce, rank-0 broadcast of the initialisation, per-step
    all-gather of updated rows (or, with Adam, the all-reduce of the zero-padded gradient, affinity_matcher.py:395-413).
    Every rank must end with the same embedding, and it must be the single-process run's."""
    import fake_ops
    from helpers import negative_table

    import torchdr_b200 as tb

    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 31000 + 2 * (os.getpid() % 1000) + scenario
    procs = [ctx.Process(target=_sharded_umap_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = dict(q.get(timeout=300) for _ in procs)
    for p in procs:
        p.join(timeout=60)
    fake_ops.install(monkeypatch)
    g = golden("umap_n300_d16_k15")
    seed, n = int(g["seed"]), 299

    class Injected(tb.UMAP):
        def on_training_step_start(self):
            self.neg_indices_ = negative_table(seed, int(self.n_iter_), n, 75)

    opt = {} if port % 2 == 0 else {"optimizer": "Adam", "lr": 0.05}
    Z1 = Injected(n_neighbors=15, max_iter=4, init=t(g["Zinit"])[:n], random_state=0, process_duplicates=False,
                  min_grad_norm=0.0, check_interval=2, distributed=False, **opt).fit_transform(t(g["X"])[:n].contiguous())
    Zs = [torch.frombuffer(bytearray(res[r]), dtype=torch.float32).reshape(n, 2) for r in range(2)]
    assert torch.equal(Zs[0], Zs[1])  # every rank holds the same embedding
    # the stand-in's row-chunked einsum sums in a different order than the full one (the CUDA kernels are bit-identical
    # across partitions, scripts/dist_check.py); 4 steps keep that noise far from the loop's amplification
    from helpers import rel_fro

    assert rel_fro(Zs[0], Z1) < 1e-5, rel_fro(Zs[0], Z1)
