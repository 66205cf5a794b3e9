"""Estimator seam: ``UMAP``, ``LargeVis``, ``TSNE``, ``InfoTSNE``, ``SNE`` with the reference's sklearn-style surface.

Mirrors ``torchdr/base.py:27-229`` (``DRModule.fit / fit_transform / transform``),
``torchdr/affinity_matcher.py:201-352`` (affinity -> init -> optimisation loop with the
``on_*`` lifecycle hooks) and ``torchdr/neighbor_embedding/{base,umap,largevis,tsne,infotsne,sne}.py``.
UMAP: the iterations between two convergence checks run in ONE persistent kernel launch (row-sharded: with the NVLink
row exchange and the cross-GPU barrier inside it); LargeVis: one row-local step per iteration (gradient gathered from
P + P^T, fused momentum SGD, rows stored into the peers); t-SNE / InfoTSNE / SNE: gradient kernel + momentum-SGD kernel.
The optimiser / scheduler objects are the reference's own ``torch.optim`` classes stepped on a dummy parameter, so
learning-rate and momentum sequences — including the reference's param-group reuse when the optimiser is rebuilt after
early exaggeration — are identical.  Inputs whose row order carries no locality are fitted in a Voronoi-tree order
(``reorder.py``) and the permutation is undone on the embedding.
"""

import logging
import os
import random
import warnings

import numpy as np
import torch
import torch.distributed as dist
import torch.nn as nn
from sklearn.base import BaseEstimator

from . import _lib, ops
from .affinity import EntropicAffinity, UMAPAffinity
from .distance import _to_device_tensor
from .distributed import PeerEmbedding, all_bounds, all_gather_rows, is_distributed, upload_sharded


def find_ab_params(spread, min_dist):
    """``torchdr/neighbor_embedding/umap.py:19-36`` (same scipy call, same grid)."""
    from scipy.optimize import curve_fit

    def curve(x, a, b):
        return 1.0 / (1.0 + a * x ** (2 * b))

    xv = np.linspace(0, spread * 3, 300)
    yv = np.zeros(xv.shape)
    yv[xv < min_dist] = 1.0
    yv[xv >= min_dist] = np.exp(-(xv[xv >= min_dist] - min_dist) / spread)
    params, _ = curve_fit(curve, xv, yv)
    return params[0].item(), params[1].item()


def _may_have_duplicate_rows(X, chunk=131072):
    """False only if all rows of X are pairwise distinct.  The reference finds duplicates with ``torch.unique(X, dim=0)``
    (base.py:132-146), a lexicographic sort of all rows that costs more than the whole fit at 1 M x 128; here every row
    gets a 64-bit multiplicative hash of its bit pattern (-0.0 canonicalised to +0.0, which ``unique`` treats as equal;
    the input is already checked finite), the hashes are sorted, and equal neighbours mean "maybe": equal rows always
    collide, so distinct hashes prove distinct rows, and on any collision the caller runs the reference's
    ``torch.unique`` unchanged."""
    n, d = X.shape
    if n < 2:
        return False
    g = torch.Generator().manual_seed(0x5DEECE66D)
    w = (torch.randint(0, 2**62, (d,), generator=g, dtype=torch.int64) * 2 + 1).to(X.device)  # odd multipliers
    h = torch.empty(n, dtype=torch.int64, device=X.device)
    for a in range(0, n, chunk):
        bits = (X[a:a + chunk] + 0.0).contiguous().view(torch.int32).to(torch.int64)
        h[a:a + chunk] = (bits * w).sum(1)  # wraps modulo 2^64
    hs = torch.sort(h).values
    return bool((hs[1:] == hs[:-1]).any())


class _NeighborEmbeddingB200(BaseEstimator, nn.Module):
    """Driver shared by the estimators (base.py:27-79 ``DRModule(BaseEstimator, nn.Module, ABC)``, affinity_matcher.py,
    neighbor_embedding/base.py).  Like the reference's classes these are sklearn estimators (``get_params`` /
    ``set_params`` / ``clone`` read the constructor signature) and torch modules (the input affinity is a sub-module;
    ``affinity_in_``, ``NN_indices_``, ``neg_indices_`` ... are non-persistent buffers dropped by ``clear_memory``)."""

    _use_closed_form_gradients = False

    def __init__(self, n_components=2, lr=1.0, optimizer="SGD", optimizer_kwargs="auto", scheduler=None,
                 scheduler_kwargs="auto", min_grad_norm=1e-7, max_iter=2000, init="pca", init_scaling=1e-4,
                 device="auto", backend=None, verbose=False, random_state=None, early_exaggeration_coeff=None,
                 early_exaggeration_iter=None, repulsion_strength=1.0, check_interval=50, compile=False,
                 distributed="auto", process_duplicates=True, precise=False, knn_order="auto", **kwargs):
        nn.Module.__init__(self)
        if n_components != 2:
            raise NotImplementedError("[TorchDR-B200] the step kernels are specialised for n_components=2.")
        if "learning_rate" in kwargs:  # NE base.py:170-171
            lr = kwargs.pop("learning_rate")
        if "early_exaggeration" in kwargs:
            early_exaggeration_coeff = kwargs.pop("early_exaggeration")
        self.n_components = n_components
        self.lr = lr
        self.optimizer = optimizer
        self.optimizer_kwargs = optimizer_kwargs
        self.scheduler = scheduler
        self.min_grad_norm = min_grad_norm
        self.max_iter = max_iter
        self.init = init
        self.init_scaling = init_scaling
        self.device = device if device is not None else "auto"
        self.backend = backend
        self.verbose = verbose
        self.random_state = random_state
        self.early_exaggeration_iter = early_exaggeration_iter or 0  # NE base.py:160-162
        self.early_exaggeration_coeff = 1 if early_exaggeration_coeff is None else early_exaggeration_coeff
        self.repulsion_strength = repulsion_strength
        self.check_interval = check_interval
        self.compile = compile
        self.process_duplicates = process_duplicates
        self.precise = precise  # fp64 transcendental evaluation in the step kernel (parity runs)
        # engine-native (not in the reference): "auto" runs the fit in a locality-creating row order when the input
        # order has none (torchdr_b200/reorder.py) and undoes the permutation on the embedding; "input" never does
        self.knn_order = knn_order
        # NE base.py:175-182 — LinearLR "auto" goes from 1 to 0 over max_iter
        if scheduler == "LinearLR" and scheduler_kwargs == "auto":
            scheduler_kwargs = {"start_factor": torch.tensor(1.0), "end_factor": torch.tensor(0), "total_iters": max_iter}
        self.scheduler_kwargs = scheduler_kwargs
        self.logger = logging.getLogger(f"torchdr_b200.{self.__class__.__name__}")
        if verbose:
            self.logger.setLevel(logging.INFO)
        if random_state is not None:  # base.py:75-79, utils/utils.py:51-97
            seed = int(random_state)
            random.seed(seed)
            np.random.seed(seed)
            torch.manual_seed(seed)
            self._actual_seed = seed
        self._setup_distributed(distributed)
        self.embedding_ = None
        self.is_fitted_ = False
        self.n_iter_ = torch.tensor(-1, dtype=torch.long)

    # ---- distributed (NE base.py:354-383) -------------------------------------------------
    def _setup_distributed(self, distributed):
        self.distributed = is_distributed() if distributed == "auto" else bool(distributed)
        if self.distributed:
            if not is_distributed():
                raise RuntimeError(
                    "[TorchDR] distributed=True requires launching with torchrun. "
                    "Example: torchrun --nproc_per_node=4 your_script.py"
                )
            self.rank, self.world_size = dist.get_rank(), dist.get_world_size()
            self.is_multi_gpu = self.world_size > 1
            if self.device == "cpu":
                raise ValueError("[TorchDR] Distributed mode requires GPU (device cannot be 'cpu')")
            local_rank = int(os.environ.get("LOCAL_RANK", 0))
            torch.cuda.set_device(local_rank)
            self.device = torch.device(f"cuda:{local_rank}")
        else:
            self.rank, self.world_size, self.is_multi_gpu = 0, 1, False

    # ---- public API (base.py:85-187) ------------------------------------------------------
    def fit(self, X, y=None):
        self.fit_transform(X, y=y)
        return self

    def _tick(self, name):
        """Stage timer (TDR_TIMING=1): synchronises and records seconds since the previous tick in ``timings_``."""
        if not self._timing:
            return
        import time

        torch.cuda.synchronize()
        now = time.perf_counter()
        self.timings_[name] = self.timings_.get(name, 0.0) + now - self._t_last
        self._t_last = now

    def fit_transform(self, X, y=None):
        self._timing = os.environ.get("TDR_TIMING") == "1"
        if self._timing:
            import time

            self.timings_, self._t_last = {}, time.perf_counter()
        was_numpy = isinstance(X, np.ndarray)
        in_device = X.device if isinstance(X, torch.Tensor) else None
        host_input = was_numpy or (isinstance(X, torch.Tensor) and X.device.type == "cpu")
        if self.world_size > 1 and host_input and getattr(X, "ndim", 0) == 2:
            # row-sharded fit on host input: each rank uploads its own row chunk, NVLink all-gather builds the database
            Xd = upload_sharded(torch.from_numpy(X) if was_numpy else X, self.device)
        else:
            Xd = _to_device_tensor(X, self.device)
            if Xd.dtype != torch.float32:
                Xd = Xd.float()
            Xd = Xd.contiguous()
        self._tick("h2d")
        if not bool(torch.isfinite(Xd).all()):
            raise ValueError("[TorchDR] ERROR : input contains NaN or infinite values.")
        self._tick("finite_check")
        if self.process_duplicates and _may_have_duplicate_rows(Xd):  # base.py:132-146
            Xu, inverse = torch.unique(Xd, dim=0, return_inverse=True)
            if Xu.shape[0] < Xd.shape[0]:
                self.logger.info(f"Detected {Xd.shape[0] - Xu.shape[0]} duplicate samples, performing DR on unique data.")
                self.embedding_ = self._fit_transform(Xu.contiguous())[inverse]
            else:
                self.embedding_ = self._fit_transform(Xd)
        else:
            self.embedding_ = self._fit_transform(Xd)
        self.is_fitted_ = True
        out = self.embedding_
        if was_numpy:
            out = out.detach().cpu().numpy()  # (a pinned staging buffer costs more to allocate than it saves: measured)
            self._tick("d2h")
            return out
        if in_device is not None and in_device != out.device:
            return out.to(in_device)
        return out

    def transform(self, X=None):
        if not self.is_fitted_:
            raise ValueError("This DRModule instance is not fitted yet. Call 'fit' or 'fit_transform' with some data first.")
        if X is not None:
            raise NotImplementedError("Transforming new data is not implemented for this model.")
        return self.embedding_

    # ---- lifecycle hooks (affinity_matcher.py:475-489) -------------------------------------
    def on_affinity_computation_start(self):
        pass

    def on_affinity_computation_end(self):
        pass

    def on_training_step_start(self):
        """Subclasses / tests may set ``self.neg_indices_`` (int64 [n_local, n_negatives], already
        adjusted as in NE base.py:629-636); when left ``None`` negatives are drawn in-kernel.

        ``discard_NNs=True``: the reference's own draw (NE base.py:638-647) — ``randint(1, N - width)`` shifted by
        ``searchsorted`` into the sorted per-row exclusion table (self + neighbours) — made with the same torch calls
        on the device and handed to the kernels as an injected table.  A functional path, not a fast one: like the
        reference it keeps an [n_local, 1 + W] table and writes n_local x n_negatives indices per step."""
        self.neg_indices_ = None
        excl = getattr(self, "negative_exclusion_indices_", None)
        if excl is not None:
            n_local = excl.shape[0]
            raw = torch.randint(1, self.n_samples_in_ - excl.shape[1], (n_local, self.n_negatives), device=excl.device)
            self.neg_indices_ = (raw + torch.searchsorted(excl, raw, right=True)).contiguous()

    def _build_negative_exclusions(self, nn_rows):
        """NE base.py:578-615: sorted [self, neighbours] per local row (the -1 padding of a symmetrised graph sorts
        first, as it does in the reference)."""
        me = self.chunk_indices_.unsqueeze(1)
        excl = torch.cat([me, nn_rows.long()], dim=1).sort(dim=1).values
        n_possible = self.n_samples_in_ - excl.shape[1]
        if self.n_negatives > n_possible and self.verbose:
            raise ValueError(f"[TorchDR] ERROR : requested {self.n_negatives} negatives but only {n_possible} available.")
        self._set_buffer("negative_exclusion_indices_", excl)  # NE base.py:605

    def on_training_step_end(self):
        pass

    # ---- checks ----------------------------------------------------------------------------
    def _check_n_neighbors(self, n):
        for name in ("perplexity", "n_neighbors"):  # NE base.py:258-267
            if hasattr(self, name) and n <= getattr(self, name):
                raise ValueError(
                    f"[TorchDR] ERROR : Number of samples is smaller than {name} ({n} <= {getattr(self, name)})."
                )

    # ---- init (affinity_matcher.py:493-573, NE base.py:410-423) ----------------------------
    def _init_embedding(self, X):
        n = X.shape[0]
        dev = X.device
        perm = getattr(self, "_perm", None)  # the fit runs on X[perm]: row r of Z belongs to original point perm[r]
        if isinstance(self.init, (torch.Tensor, np.ndarray)):
            Z = torch.as_tensor(self.init).to(device=dev, dtype=torch.float32)
            if perm is not None:
                Z = Z[perm]
        elif self.init in ("normal", "random"):
            Z = torch.randn((n, self.n_components), device=dev, dtype=torch.float32)
            if perm is not None:
                Z = Z[perm]  # every original point keeps the draw it would get in the input order
        elif self.init == "pca":
            Z = _pca_init(X, self.n_components)
        else:
            raise ValueError(f"[TorchDR] ERROR : init {self.init} not supported in {self.__class__.__name__}.")
        Z = (self.init_scaling * Z / Z[:, 0].std()).contiguous()
        if self.world_size > 1:
            dist.broadcast(Z, src=0)
        return Z

    # ---- optimiser / scheduler: the reference's torch.optim objects on a dummy parameter ----
    def _set_learning_rate(self):
        if self.lr == "auto":  # NE base.py:299-310
            if self.optimizer != "SGD" and self.verbose:
                warnings.warn("[TorchDR] WARNING : when 'auto' is used for the learning rate, the optimizer should be 'SGD'.")
            self.lr_ = max(self.n_samples_in_ / self.early_exaggeration_coeff_ / 4, 50)
        else:
            self.lr_ = self.lr

    def _optimizer_spec(self):
        """(class, kwargs) exactly as NE base.py:312-343 resolves them."""
        if isinstance(self.optimizer, str):
            cls = getattr(torch.optim, self.optimizer, None)
            if cls is None:
                raise ValueError(f"[TorchDR] ERROR: Optimizer '{self.optimizer}' not found in torch.optim")
        else:
            if not (isinstance(self.optimizer, type) and issubclass(self.optimizer, torch.optim.Optimizer)):
                raise ValueError("[TorchDR] ERROR: optimizer must be a string (name of an optimizer in "
                                 "torch.optim) or a subclass of torch.optim.Optimizer")
            cls = self.optimizer
        if self.optimizer_kwargs == "auto":  # NE base.py:331-338
            if self.optimizer == "SGD":
                kw = {"momentum": 0.5} if self.early_exaggeration_coeff_ > 1 else {"momentum": 0.8}
            else:
                kw = {}
        else:
            kw = self.optimizer_kwargs or {}
        return cls, kw

    def _uses_native_sgd(self):
        """The update kernels implement torch.optim.SGD with (only) a momentum option.  Every other optimiser or
        option runs as the reference runs it: the torch optimiser itself, stepped on the embedding with the
        gradient the kernels computed (torch library code on the device: plumbing around the gradient kernels)."""
        cls, kw = self._optimizer_spec()
        return cls is torch.optim.SGD and set(kw) <= {"momentum"}

    def _configure_optimizer(self):
        cls, kw = self._optimizer_spec()
        self.optimizer_ = cls(self.params_, lr=self.lr_, **kw)  # NE base.py:343
        return self.optimizer_

    def _configure_scheduler(self):
        if self.early_exaggeration_coeff_ > 1:  # NE base.py:345-350
            n_iter = min(self.early_exaggeration_iter, self.max_iter)
        else:
            n_iter = self.max_iter - self.early_exaggeration_iter
        del n_iter  # affinity_matcher.py:620-657 never forwards it to the scheduler class
        if self.scheduler is None:
            self.scheduler_ = None
        elif isinstance(self.scheduler, str):
            cls = getattr(torch.optim.lr_scheduler, self.scheduler, None)
            if cls is None:
                raise ValueError(f"[TorchDR] ERROR: Scheduler '{self.scheduler}' not found in torch.optim.lr_scheduler.")
            self.scheduler_ = cls(self.optimizer_, **(self.scheduler_kwargs or {}))
        else:
            self.scheduler_ = self.scheduler(self.optimizer_, **(self.scheduler_kwargs or {}))
        return self.scheduler_

    def _generic_step(self, grad, check):
        """affinity_matcher.py:414-429 with the reference's own optimiser object: grad -> optimizer_.step()."""
        self._dummy.grad = grad
        self.optimizer_.step()
        if self.scheduler_ is not None:
            self.scheduler_.step()
        if check:
            self._gnorm.copy_(grad.double().pow(2).sum().reshape(1))
            self._nan.copy_(torch.isnan(self._dummy.detach()).any().to(torch.int32).reshape(1))

    def _hyper(self):
        g = self.optimizer_.param_groups[0]
        return float(g["lr"]), float(g.get("momentum", 0.0))

    def _advance_schedule(self):
        self.optimizer_.step()  # dummy parameter has no grad: keeps torch's step-order bookkeeping quiet
        if self.scheduler_ is not None:
            self.scheduler_.step()

    # ---- main driver (affinity_matcher.py:201-352) -----------------------------------------
    def _fit_order(self, X):
        """Row order of this fit: None (the input's) or a permutation (rank 0's, broadcast, when row-sharded).  Subclass
        hooks and injected tables see row indices, so they pin the input order."""
        base = _NeighborEmbeddingB200
        user_hooks = any(getattr(type(self), h) is not getattr(base, h) for h in
                         ("on_training_step_start", "on_training_step_end", "on_affinity_computation_start",
                          "on_affinity_computation_end"))
        if self.knn_order == "input" or user_hooks or self.precise:
            return None
        from .reorder import choose_order

        if self.world_size == 1:
            return choose_order(X, self.knn_order)
        flag = torch.zeros(1, dtype=torch.int64, device=X.device)
        perm = None
        if self.rank == 0:
            perm = choose_order(X, self.knn_order)
            flag.fill_(0 if perm is None else 1)
        dist.broadcast(flag, src=0)
        if int(flag.item()) == 0:
            return None
        if perm is None:
            perm = torch.empty(X.shape[0], dtype=torch.int64, device=X.device)
        dist.broadcast(perm, src=0)
        return perm

    def _fit_transform(self, X):
        self._perm = self._fit_order(X)
        if self._perm is not None:
            X = X[self._perm].contiguous()
            self.affinity_in.knn_order = "presorted"
        else:
            self.affinity_in.knn_order = "input"
        self._tick("order")
        Z = self._fit_transform_ordered(X)
        if self._perm is not None:
            out = torch.empty_like(Z)
            out[self._perm] = Z
            self.embedding_ = Z = out
        self._perm = None
        return Z

    def _fit_transform_ordered(self, X):
        n = X.shape[0]
        self._check_n_neighbors(n)
        self.early_exaggeration_coeff_ = self.early_exaggeration_coeff  # NE base.py:276
        self.n_samples_in_, self.n_features_in_ = X.shape
        self.device_ = X.device
        _lib.require_device(X.device)
        bounds = all_bounds(n, self.world_size)
        self._bounds = bounds
        self.chunk_start_, self.chunk_end_ = bounds[self.rank]

        self.on_affinity_computation_start()
        if self.verbose:
            self.logger.info(f"----- Computing the input affinity matrix with {self.affinity_in.__class__.__name__} -----")
        self._compute_affinity(X)
        self._tick("affinity+graph")
        self._set_buffer("chunk_indices_", torch.arange(self.chunk_start_, self.chunk_end_, device=X.device))  # NE base.py:406-408
        if getattr(self, "discard_NNs", False):
            self._build_negative_exclusions(self._neighbour_rows())
        self.on_affinity_computation_end()

        if self.verbose:
            self.logger.info("----- Optimizing the embedding -----")
        Z = self._init_embedding(X)
        self._native_opt = self._uses_native_sgd()
        # native: the torch objects only keep the books (lr / momentum sequence) on a dummy parameter; generic: they
        # own the embedding (same storage as Z).  ONE dict reused by every rebuild (affinity_matcher.py:588-590).
        self._dummy = torch.nn.Parameter(torch.zeros(1) if self._native_opt else Z)
        self.params_ = [{"params": [self._dummy]}]
        self._set_learning_rate()
        self._configure_optimizer()
        self._configure_scheduler()
        del X

        dev = Z.device
        self._gnorm = torch.zeros(1, dtype=torch.float64, device=dev)
        self._nan = torch.zeros(1, dtype=torch.int32, device=dev)
        self.embedding_ = Z
        self._tick("init")
        self._loop()
        self._tick("loop")
        self.n_iter_ = torch.tensor(self._last_step, dtype=torch.long)
        self.clear_memory()
        return self.embedding_

    def _kernel_seed(self):
        """Key of the in-kernel negative stream: the estimator's seed, else torch's current seed (unseeded runs differ
        from process to process like the reference's torch.randint draws, NE base.py:629)."""
        if self.random_state is not None:
            return int(self._actual_seed)
        return int(torch.initial_seed() % (2**63))

    def _check_nan(self, step):
        """check_NaNs of affinity_matcher.py:315-319.  Row-sharded runs agree on the flag first (MAX over ranks): a
        rank that alone saw a NaN must not leave its peers waiting in the next iteration's exchange."""
        if self.world_size > 1:
            dist.all_reduce(self._nan, op=dist.ReduceOp.MAX)
        if int(self._nan.item()):
            raise ValueError(f"[TorchDR] ERROR AffinityMatcher : NaNs in the embeddings at iter {step}.")

    def _converged(self, step, grad_norm):
        if self.verbose:
            lr, _ = self._hyper()
            self.logger.info(f"[{step}/{self.max_iter}] Grad norm: {grad_norm:.2e} | LR: {lr:.2e}")
        if grad_norm < self.min_grad_norm:  # affinity_matcher.py:343-349
            if self.verbose:
                self.logger.info(f"Convergence reached at iter {step} with grad norm: {grad_norm:.2e}.")
            return True
        return False

    def _set_buffer(self, name, tensor):
        """register_buffer(..., persistent=False) as in affinity_matcher.py:269-286 / NE base.py:605,649 (re-entrant)."""
        if name in self._buffers:
            self._buffers[name] = tensor
        else:
            if name in self.__dict__:
                del self.__dict__[name]
            self.register_buffer(name, tensor, persistent=False)

    def clear_memory(self):
        """affinity_matcher.py:661-677: drop every non-persistent buffer (and the loop's scratch state)."""
        for name in list(self._non_persistent_buffers_set):
            if hasattr(self, name):
                delattr(self, name)
        for name in ("_gnorm", "_nan", "optimizer_", "scheduler_", "params_", "_dummy", "neg_indices_", "_graph",
                     "affinity_in_", "NN_indices_", "chunk_indices_", "_mom", "_grad", "negative_exclusion_indices_"):
            if hasattr(self, name):
                delattr(self, name)
        if hasattr(self.affinity_in, "clear_memory"):
            self.affinity_in.clear_memory()

    # ---- autograd-mode loop shared by LargeVis and TSNE (gradient kernel + momentum SGD) ----
    def _loop(self):
        Z = self.embedding_
        n = Z.shape[0]
        self._grad = torch.zeros_like(Z)
        self._mom = torch.zeros_like(Z)
        first = True
        self._last_step = -1
        for step in range(self.max_iter):
            self.n_iter_ = torch.tensor(step, dtype=torch.long)
            self._last_step = step
            self.on_training_step_start()
            check = step % self.check_interval == 0
            self._grad.zero_()
            self._compute_gradient(Z, step)
            if self.world_size > 1:  # affinity_matcher.py:424-425
                dist.all_reduce(self._grad, op=dist.ReduceOp.SUM)
            if self._native_opt:
                lr, mom = self._hyper()
                if check:
                    self._gnorm.zero_()
                ops.sgd_momentum(Z, self._mom, self._grad, lr, mom, first or mom == 0.0,
                                 gnorm_sq=self._gnorm if check else None, nan_flag=self._nan)
                first = False
                self._advance_schedule()
            else:
                self._generic_step(self._grad, check)
            # NE base.py:282-295 — end of early exaggeration: rebuild optimiser (+ scheduler)
            if self.early_exaggeration_coeff_ > 1 and step == self.early_exaggeration_iter:
                self.early_exaggeration_coeff_ = 1
                self._set_learning_rate()
                self._configure_optimizer()
                self._configure_scheduler()
                first = True  # fresh optimiser state: momentum buffer restarts
            self.on_training_step_end()
            if check:
                self._check_nan(step)
                if self._converged(step, float(self._gnorm.item()) ** 0.5):
                    break
        self._check_nan(self._last_step)


def _pca_init(X, q):
    """PCA scores for ``init="pca"`` via covariance + eigh on the device (one-off; torch library
    calls: a d x d GEMM and a d x d eigh).  The reference uses a full SVD
    (spectral_embedding/pca.py:171-178); both give the top-q principal scores up to sign, and the
    sign convention below follows svd_flip(u_based_decision) of utils/utils.py:264-301."""
    mean = X.mean(0, keepdim=True)
    Xc = X - mean
    cov = (Xc.T @ Xc).double()
    w, V = torch.linalg.eigh(cov)
    comp = V[:, -q:].flip(1).float()  # [d, q], descending eigenvalue
    scores = Xc @ comp
    amax = scores.abs().argmax(0)
    signs = torch.sign(scores[amax, torch.arange(q, device=X.device)])
    signs[signs == 0] = 1
    return scores * signs


class UMAP(_NeighborEmbeddingB200):
    """``torchdr/neighbor_embedding/umap.py:39-292``."""

    _use_closed_form_gradients = True

    def __init__(self, n_neighbors=30, n_components=2, min_dist=0.1, spread=1.0, a=None, b=None, lr=1e0,
                 optimizer="SGD", optimizer_kwargs=None, scheduler="LinearLR", scheduler_kwargs="auto", init="pca",
                 init_scaling=1e-4, min_grad_norm=1e-7, max_iter=1000, device="auto", backend=None, verbose=False,
                 random_state=None, max_iter_affinity=100, metric="sqeuclidean", negative_sample_rate=5,
                 check_interval=50, discard_NNs=False, compile=False, distributed="auto", **kwargs):
        self.n_neighbors = n_neighbors
        self.min_dist = min_dist
        self.spread = spread
        self.metric = metric
        self.max_iter_affinity = max_iter_affinity
        self.negative_sample_rate = negative_sample_rate
        self.sparsity = True
        self._eps = 1e-3
        self.a, self.b = a, b  # as passed: sklearn's get_params / clone read them (the reference raises here)
        if a is None or b is None:
            a, b = find_ab_params(spread, min_dist)
        self._a, self._b = a, b
        self.n_negatives = int(negative_sample_rate * n_neighbors)  # umap.py:177
        self.discard_NNs = discard_NNs
        super().__init__(n_components=n_components, lr=lr, optimizer=optimizer, optimizer_kwargs=optimizer_kwargs,
                         scheduler=scheduler, scheduler_kwargs=scheduler_kwargs, min_grad_norm=min_grad_norm,
                         max_iter=max_iter, init=init, init_scaling=init_scaling, device=device, backend=backend,
                         verbose=verbose, random_state=random_state, check_interval=check_interval, compile=compile,
                         distributed=distributed, **kwargs)
        self.affinity_in = UMAPAffinity(n_neighbors=n_neighbors, metric=metric, max_iter=max_iter_affinity,
                                        device=self.device, backend=backend, verbose=verbose, sparsity=True,
                                        compile=compile, distributed=self.distributed)

    def _compute_affinity(self, X):
        rowptr, col, val = self.affinity_in.compute_csr(X)
        self._tick("knn+sigma+symmetrise")
        # umap.py:215-234 — threshold A_max/max_iter, epochs_per_sample, epoch_of_next_sample
        a_max = ops.max_value(val)
        if self.world_size > 1:
            dist.all_reduce(a_max, op=dist.ReduceOp.MAX)
        a_max = float(a_max.item())
        eps, _ = ops.umap_schedule(val, a_max, self.max_iter)
        if self.verbose:
            kept = float((eps < float("inf")).float().mean().item()) * 100
            self.logger.info(f"Keeping {kept:.1f}% of affinity edges.")
        self._graph_full = (rowptr, col, val, eps)
        self._graph = ops.umap_compact(rowptr, col, eps)  # (rowptr, col, eps, eons) of live edges

    def _neighbour_rows(self):
        # the reference's NN_indices_ of UMAP: the symmetrised graph's padded index matrix (affinity/base.py:407-431)
        rowptr, col, val, _ = self._graph_full
        return ops.csr_to_ell(rowptr, col, val)[1]

    def clear_memory(self):
        if hasattr(self, "_graph_full"):
            del self._graph_full
        super().clear_memory()

    def _loop(self):
        rowptr, col, eps, eons = self._graph
        s, e = self.chunk_start_, self.chunk_end_
        Za = self.embedding_
        seed = self._kernel_seed()
        rep = float(self.repulsion_strength)
        if not self._native_opt or self._hyper()[1] != 0.0:
            # the step kernel fuses plain SGD (umap.py:139, the default); anything else takes the gradient from the
            # kernel and lets the reference's optimiser object apply it (affinity_matcher.py:395-413)
            return self._loop_generic_optimizer(seed, rep)
        self._last_step = -1
        step = 0
        stop = False
        hooks_per_step = type(self).on_training_step_start is not UMAP.on_training_step_start or \
            type(self).on_training_step_end is not _NeighborEmbeddingB200.on_training_step_end or \
            getattr(self, "negative_exclusion_indices_", None) is not None
        # Batched branches run a whole check_interval of iterations in ONE persistent kernel launch
        # (tdr_umap_run_f32 / tdr_umap_run_p2p_f32); row-sharded, the kernel also stores the updated rows into every
        # peer over NVLink and runs the per-iteration barrier itself (PeerEmbedding).  Per-step branch: hooks that
        # must run between iterations, injected negatives, the fp64 parity kernel, or no symmetric memory (then one
        # NCCL all-gather per iteration).
        peer, cur = None, 0
        self.exchange_ = "none" if self.world_size == 1 else "nccl-allgather"
        if self.world_size > 1 and not hooks_per_step and not self.precise and os.environ.get("TDR_NO_P2P") != "1":
            try:
                peer = PeerEmbedding.get(Za)
                Za, Zb = peer.bufs[0], peer.bufs[1]
                self.exchange_ = "p2p-fused"
            except Exception as exc:  # symmetric memory unavailable on this system
                self.logger.warning(f"symmetric memory unavailable ({exc}); falling back to NCCL all-gather")
                peer = None
        if peer is None:
            pair = torch.empty((2,) + tuple(Za.shape), dtype=Za.dtype, device=Za.device)  # the two buffers side by side
            pair[0], pair[1] = Za, Za
            Za, Zb = pair[0], pair[1]
        sync = peer.sync if peer is not None else ops.RunSync(Za.device)
        # Learning rates come from the reference's own optimizer / scheduler objects (~25 us of host time per step).
        # In the batched branches they are produced one batch AHEAD, after the current batch has been launched and
        # before its convergence check synchronises, so the bookkeeping overlaps the kernels instead of idling the GPU.
        # A batch that never runs leaves the objects advanced: harmless, they are discarded after the fit.
        lr_buf = []  # lr_buf[t] = learning rate of step t

        def lrs_for(a, b):
            while len(lr_buf) <= b:
                lr_buf.append(self._hyper()[0])
                self._advance_schedule()
            return lr_buf[a:b + 1]

        while step < self.max_iter and not stop:
            lam = float(self.early_exaggeration_coeff_)
            # batch = steps up to (and including) the next one with n_iter % check_interval == 0 — or the last step
            # of early exaggeration, after which the optimiser is rebuilt (NE base.py:282-295)
            nxt_check = step if step % self.check_interval == 0 else (step // self.check_interval + 1) * self.check_interval
            last = min(nxt_check, self.max_iter - 1)
            exag_end = self.early_exaggeration_coeff_ > 1 and step <= self.early_exaggeration_iter <= last
            if exag_end:
                last = self.early_exaggeration_iter
            ahead = min(last + self.check_interval, self.max_iter - 1)
            if self.early_exaggeration_coeff_ > 1 and last < self.early_exaggeration_iter:
                ahead = min(ahead, self.early_exaggeration_iter)  # the optimiser is rebuilt there: no look-ahead past it
            want = last % self.check_interval == 0
            if want:
                self._gnorm.zero_()
            if peer is not None or not (self.world_size > 1 or hooks_per_step):
                lrs = lrs_for(step, last)
                if peer is not None:
                    cur = ops.umap_run_p2p(peer, cur, s, e - s, rowptr, col, eps, eons, step, lrs, self._a, self._b,
                                           n_neg=self.n_negatives, rate=self.negative_sample_rate, seed=seed, lam=lam,
                                           repulsion=rep, gnorm_sq=self._gnorm if want else None, nan_flag=self._nan)
                    Za, Zb = peer.bufs[cur], peer.bufs[1 - cur]
                else:
                    res = ops.umap_run(Za, Zb, rowptr, col, eps, eons, step, lrs, self._a, self._b,
                                       n_neg=self.n_negatives, rate=self.negative_sample_rate, seed=seed, lam=lam,
                                       repulsion=rep, precise=self.precise, gnorm_sq=self._gnorm if want else None,
                                       nan_flag=self._nan, sync=sync)
                    if res is not Za:
                        Za, Zb = Zb, Za
                self.n_iter_ = torch.tensor(last, dtype=torch.long)
                self.embedding_ = Za
                if not exag_end:
                    lrs_for(last + 1, ahead)
            else:
                for t in range(step, last + 1):
                    self.n_iter_ = torch.tensor(t, dtype=torch.long)
                    self.on_training_step_start()
                    lr = self._hyper()[0]
                    neg = getattr(self, "neg_indices_", None)
                    ops.umap_step(Za, Zb, s, e - s, rowptr, col, eps, eons, t, self._a, self._b, lr,
                                  neg=neg, n_neg=self.n_negatives, rate=self.negative_sample_rate, seed=seed, lam=lam,
                                  repulsion=rep, precise=self.precise,
                                  gnorm_sq=self._gnorm if (want and t == last) else None, nan_flag=self._nan)
                    if self.world_size > 1:
                        all_gather_rows(Zb, self._bounds, self.rank)
                    Za, Zb = Zb, Za
                    self.embedding_ = Za
                    self._advance_schedule()
                    self.on_training_step_end()
            if want and self.world_size > 1:
                dist.all_reduce(self._gnorm, op=dist.ReduceOp.SUM)
            self._last_step = last
            step = last + 1
            if peer is not None or self.world_size == 1:
                try:
                    sync.check()
                except _lib.B200EngineError:
                    PeerEmbedding._cache.clear()  # an aborted exchange leaves flags / epochs undefined: never reuse them
                    raise
            self._check_nan(last)
            if exag_end:  # NE base.py:282-295: coefficient back to 1, fresh optimiser and scheduler
                self.early_exaggeration_coeff_ = 1
                self._set_learning_rate()
                self._configure_optimizer()
                self._configure_scheduler()
                del lr_buf[step:]
            if want:
                stop = self._converged(last, float(self._gnorm.item()) ** 0.5)
        self.embedding_ = Za.clone() if peer is not None else Za  # leave symmetric memory: the buffers are reused


def _umap_loop_generic_optimizer(self, seed, rep):
    """UMAP iterations with an arbitrary torch optimiser: gradient of the local rows from the step kernel
    (`grad_out`), zero-padded to N x 2 and summed over ranks as in affinity_matcher.py:395-413, then
    `optimizer_.step()` on the embedding; the kernel's own SGD output is scratch."""
    rowptr, col, eps, eons = self._graph
    s, e = self.chunk_start_, self.chunk_end_
    if self._native_opt:  # SGD with momentum: give the torch objects the real embedding
        self._native_opt = False
        self._dummy = torch.nn.Parameter(self.embedding_)
        self.params_[0]["params"] = [self._dummy]
        self._configure_optimizer()
        self._configure_scheduler()
    Z = self.embedding_
    scratch = torch.empty_like(Z)
    G = torch.zeros_like(Z)
    self._last_step = -1
    for step in range(self.max_iter):
        self.n_iter_ = torch.tensor(step, dtype=torch.long)
        self._last_step = step
        self.on_training_step_start()
        check = step % self.check_interval == 0
        if self.world_size > 1:
            G.zero_()
        ops.umap_step(Z, scratch, s, e - s, rowptr, col, eps, eons, step, self._a, self._b, 0.0,
                      neg=getattr(self, "neg_indices_", None), n_neg=self.n_negatives, rate=self.negative_sample_rate,
                      seed=seed, lam=float(self.early_exaggeration_coeff_), repulsion=rep, precise=self.precise,
                      grad_out=G[s:e])
        if self.world_size > 1:
            dist.all_reduce(G, op=dist.ReduceOp.SUM)
        self._generic_step(G, check)
        if self.early_exaggeration_coeff_ > 1 and step == self.early_exaggeration_iter:  # NE base.py:282-295
            self.early_exaggeration_coeff_ = 1
            self._set_learning_rate()
            self._configure_optimizer()
            self._configure_scheduler()
        self.on_training_step_end()
        if check:
            self._check_nan(step)
            if self._converged(step, float(self._gnorm.item()) ** 0.5):
                break
    self._check_nan(self._last_step)


UMAP._loop_generic_optimizer = _umap_loop_generic_optimizer


class _EntropicInputMixin:
    def _neighbour_rows(self):
        return self.NN_indices_

    def _compute_affinity(self, X):
        if not self.sparsity:
            # the reference's dense mode (affinity_matcher.py:280-286: an N x N affinity_in_, no NN_indices_) is outside
            # the accelerated path: the gradient kernels walk kNN rows.  EntropicAffinity(sparsity=False) itself — the
            # dense affinity of BASELINE configs[2] — is available on its own.
            raise NotImplementedError(f"[TorchDR-B200] {self.__class__.__name__}(sparsity=False): the dense N x N "
                                      "optimisation is outside the accelerated path; use sparsity=True.")
        P, idx = self.affinity_in(X, log=False, return_indices=True)
        self._set_buffer("affinity_in_", P.contiguous())      # affinity_matcher.py:269-286
        self._set_buffer("NN_indices_", idx.contiguous())


class LargeVis(_EntropicInputMixin, _NeighborEmbeddingB200):
    """``torchdr/neighbor_embedding/largevis.py:21-201``."""

    def __init__(self, perplexity=30, n_components=2, lr="auto", optimizer="SGD", optimizer_kwargs="auto",
                 scheduler="LinearLR", scheduler_kwargs=None, init="pca", init_scaling=1e-4, min_grad_norm=1e-7,
                 max_iter=1000, device="auto", backend=None, verbose=False, random_state=None, max_iter_affinity=100,
                 metric="sqeuclidean", n_negatives=5, sparsity=True, early_exaggeration_coeff=None,
                 early_exaggeration_iter=None, check_interval=50, discard_NNs=False, compile=False,
                 distributed="auto", row_local=True, **kwargs):
        self.metric = metric
        self.perplexity = perplexity
        self.max_iter_affinity = max_iter_affinity
        self.sparsity = sparsity
        self.n_negatives = n_negatives
        self.discard_NNs = discard_NNs
        self.row_local = row_local  # engine-native: False forces the reference's scatter + all-reduce formulation
        super().__init__(n_components=n_components, lr=lr, optimizer=optimizer, optimizer_kwargs=optimizer_kwargs,
                         scheduler=scheduler, scheduler_kwargs=scheduler_kwargs, min_grad_norm=min_grad_norm,
                         max_iter=max_iter, init=init, init_scaling=init_scaling, device=device, backend=backend,
                         verbose=verbose, random_state=random_state,
                         early_exaggeration_coeff=early_exaggeration_coeff,
                         early_exaggeration_iter=early_exaggeration_iter, check_interval=check_interval,
                         compile=compile, distributed=distributed, **kwargs)
        self.affinity_in = EntropicAffinity(perplexity=perplexity, metric=metric, max_iter=max_iter_affinity,
                                            device=self.device, backend=backend, verbose=verbose, sparsity=sparsity,
                                            distributed=self.distributed)

    def _row_local(self):
        """True when the fit can use the row-local step (tdr_largevis_step_f32): momentum SGD fused in the kernel,
        negatives drawn in-kernel, no per-step hooks.  Otherwise: scatter kernel + all-reduce, as the reference."""
        base = _NeighborEmbeddingB200
        hooks = any(getattr(type(self), h) is not getattr(base, h) for h in ("on_training_step_start", "on_training_step_end"))
        return self.row_local and self._uses_native_sgd() and not hooks and not self.discard_NNs

    def _compute_affinity(self, X):
        super()._compute_affinity(X)
        self._graph = None
        if self._row_local():
            # union graph S = P + P^T of the local rows (the same edge exchange as UMAP's symmetrisation): row i's
            # attraction then gathers its own edges and the edges pointing at it — no scatter, no N x 2 all-reduce
            n = X.shape[0]
            s, e = self.chunk_start_, self.chunk_end_
            ext = None
            if self.world_size > 1:
                from .distributed import exchange_edges

                counts, er, ec, ev = ops.symmetrize_export(self.affinity_in_, self.NN_indices_, s, n, self.world_size, self.rank)
                ext = exchange_edges(counts, er, ec, ev)
            self._graph = ops.symmetrize_csr(self.affinity_in_, self.NN_indices_, s, n, ext=ext, mode="sum")

    def _loop(self):
        if getattr(self, "_graph", None) is None:
            return super()._loop()
        rowptr, col, val = self._graph
        s, e = self.chunk_start_, self.chunk_end_
        n_local = e - s
        Za = self.embedding_
        dev = Za.device
        peer, cur = None, 0
        self.exchange_ = "none" if self.world_size == 1 else "nccl-allgather"
        if self.world_size > 1 and os.environ.get("TDR_NO_P2P") != "1":
            try:
                peer = PeerEmbedding.get(Za)
                Za, Zb = peer.bufs[0], peer.bufs[1]
                self.exchange_ = "p2p-fused"
            except Exception as exc:  # symmetric memory unavailable on this system
                self.logger.warning(f"symmetric memory unavailable ({exc}); falling back to NCCL all-gather")
                peer = None
        if peer is None:
            pair = torch.empty((2,) + tuple(Za.shape), dtype=Za.dtype, device=dev)
            pair[0], pair[1] = Za, Za
            Za, Zb = pair[0], pair[1]
        self._mom = torch.zeros((n_local, 2), dtype=torch.float32, device=dev)
        self._grad = torch.empty((n_local, 2), dtype=torch.float32, device=dev)
        seed, rep = self._kernel_seed(), float(self.repulsion_strength)
        first = True
        self._last_step = -1
        for step in range(self.max_iter):
            self.n_iter_ = torch.tensor(step, dtype=torch.long)
            self._last_step = step
            check = step % self.check_interval == 0
            lr, mom = self._hyper()
            if check:
                self._gnorm.zero_()
            ops.largevis_step(Za, Zb, s, n_local, rowptr, col, val, self._grad, self._mom, step, lr, mom,
                              first or mom == 0.0, n_neg=self.n_negatives, seed=seed,
                              lam=float(self.early_exaggeration_coeff_), repulsion=rep,
                              gnorm_sq=self._gnorm if check else None, nan_flag=self._nan,
                              peer_ptrs=peer.peer_ptrs(1 - cur) if peer is not None else ())
            first = False
            if peer is not None:
                peer.barrier(1 - cur)  # every rank's rows have landed in every copy of the new buffer
                cur = 1 - cur
            elif self.world_size > 1:
                all_gather_rows(Zb, self._bounds, self.rank)
            Za, Zb = Zb, Za
            self.embedding_ = Za
            self._advance_schedule()
            if self.early_exaggeration_coeff_ > 1 and step == self.early_exaggeration_iter:  # NE base.py:282-295
                self.early_exaggeration_coeff_ = 1
                self._set_learning_rate()
                self._configure_optimizer()
                self._configure_scheduler()
                first = True
            if check:
                if self.world_size > 1:
                    dist.all_reduce(self._gnorm, op=dist.ReduceOp.SUM)
                self._check_nan(step)
                if self._converged(step, float(self._gnorm.item()) ** 0.5):
                    break
        self._check_nan(self._last_step)
        self.embedding_ = Za.clone() if peer is not None else Za

    def _compute_gradient(self, Z, step):
        s, e = self.chunk_start_, self.chunk_end_
        seed = self._kernel_seed()
        ops.largevis_grad(Z, s, e - s, self.affinity_in_, self.NN_indices_, self._grad, step,
                          neg=getattr(self, "neg_indices_", None), n_neg=self.n_negatives, seed=seed,
                          lam=float(self.early_exaggeration_coeff_), repulsion=float(self.repulsion_strength))


class TSNE(_EntropicInputMixin, _NeighborEmbeddingB200):
    """``torchdr/neighbor_embedding/tsne.py:22-180``."""

    def __init__(self, perplexity=30, n_components=2, lr="auto", optimizer="SGD", optimizer_kwargs="auto",
                 scheduler=None, scheduler_kwargs=None, init="pca", init_scaling=1e-4, min_grad_norm=1e-7,
                 max_iter=2000, device="auto", backend=None, verbose=False, random_state=None,
                 early_exaggeration_coeff=12.0, early_exaggeration_iter=250, max_iter_affinity=100,
                 metric="sqeuclidean", sparsity=True, check_interval=50, compile=False, distributed="auto", **kwargs):
        self.metric = metric
        self.perplexity = perplexity
        self.max_iter_affinity = max_iter_affinity
        self.sparsity = sparsity
        super().__init__(n_components=n_components, lr=lr, optimizer=optimizer, optimizer_kwargs=optimizer_kwargs,
                         scheduler=scheduler, scheduler_kwargs=scheduler_kwargs, min_grad_norm=min_grad_norm,
                         max_iter=max_iter, init=init, init_scaling=init_scaling, device=device, backend=backend,
                         verbose=verbose, random_state=random_state,
                         early_exaggeration_coeff=early_exaggeration_coeff,
                         early_exaggeration_iter=early_exaggeration_iter, check_interval=check_interval,
                         compile=compile, distributed=distributed, **kwargs)
        self.affinity_in = EntropicAffinity(perplexity=perplexity, metric=metric, max_iter=max_iter_affinity,
                                            device=self.device, backend=backend, verbose=verbose, sparsity=sparsity,
                                            distributed=self.distributed)

    def _compute_gradient(self, Z, step):
        s, e = self.chunk_start_, self.chunk_end_
        if not hasattr(self, "_tsne_ws"):
            self._tsne_ws = ops.tsne_workspace(e - s, Z.device)
        lam, rep = float(self.early_exaggeration_coeff_), float(self.repulsion_strength)
        ops.tsne_grad(Z, s, e - s, self.affinity_in_, self.NN_indices_, lam, 0, self._grad, self._tsne_ws, repulsion=rep)
        if self.world_size > 1:
            # the reference has every rank compute the whole N x N term and divide by W (tsne.py:172-180);
            # here each rank owns a row range of the double sum and the scalar normaliser is all-reduced
            S = self._tsne_ws[:8].view(torch.float64)
            dist.all_reduce(S, op=dist.ReduceOp.SUM)
        ops.tsne_grad(Z, s, e - s, self.affinity_in_, self.NN_indices_, lam, 1, self._grad, self._tsne_ws, repulsion=rep)

    def clear_memory(self):
        if hasattr(self, "_tsne_ws"):
            del self._tsne_ws
        super().clear_memory()


class InfoTSNE(_EntropicInputMixin, _NeighborEmbeddingB200):
    """``torchdr/neighbor_embedding/infotsne.py:14-197``."""

    def __init__(self, perplexity=30, n_components=2, lr="auto", optimizer="SGD", optimizer_kwargs="auto",
                 scheduler="LinearLR", scheduler_kwargs=None, init="pca", init_scaling=1e-4, min_grad_norm=1e-7,
                 max_iter=1000, device="auto", backend=None, verbose=False, random_state=None,
                 early_exaggeration_coeff=12, early_exaggeration_iter=250, max_iter_affinity=100,
                 metric="sqeuclidean", n_negatives=300, sparsity=True, check_interval=50, discard_NNs=False,
                 compile=False, distributed="auto", **kwargs):
        self.metric = metric
        self.perplexity = perplexity
        self.max_iter_affinity = max_iter_affinity
        self.sparsity = sparsity
        self.n_negatives = n_negatives
        self.discard_NNs = discard_NNs
        super().__init__(n_components=n_components, lr=lr, optimizer=optimizer, optimizer_kwargs=optimizer_kwargs,
                         scheduler=scheduler, scheduler_kwargs=scheduler_kwargs, min_grad_norm=min_grad_norm,
                         max_iter=max_iter, init=init, init_scaling=init_scaling, device=device, backend=backend,
                         verbose=verbose, random_state=random_state,
                         early_exaggeration_coeff=early_exaggeration_coeff,
                         early_exaggeration_iter=early_exaggeration_iter, check_interval=check_interval,
                         compile=compile, distributed=distributed, **kwargs)
        self.affinity_in = EntropicAffinity(perplexity=perplexity, metric=metric, max_iter=max_iter_affinity,
                                            device=self.device, backend=backend, verbose=verbose, sparsity=sparsity,
                                            distributed=self.distributed)

    def _compute_gradient(self, Z, step):
        s, e = self.chunk_start_, self.chunk_end_
        seed = self._kernel_seed()
        ops.infotsne_grad(Z, s, e - s, self.affinity_in_, self.NN_indices_, self._grad, step,
                          neg=getattr(self, "neg_indices_", None), n_neg=self.n_negatives, seed=seed,
                          lam=float(self.early_exaggeration_coeff_), repulsion=float(self.repulsion_strength))


class SNE(_EntropicInputMixin, _NeighborEmbeddingB200):
    """``torchdr/neighbor_embedding/sne.py:21-179``."""

    def __init__(self, perplexity=30, n_components=2, lr="auto", optimizer="SGD", optimizer_kwargs="auto",
                 scheduler=None, scheduler_kwargs=None, init="pca", init_scaling=1e-4, min_grad_norm=1e-7,
                 max_iter=2000, device="auto", backend=None, verbose=False, random_state=None, max_iter_affinity=100,
                 metric="sqeuclidean", sparsity=True, early_exaggeration_coeff=None, early_exaggeration_iter=None,
                 check_interval=50, compile=False, distributed="auto", **kwargs):
        self.metric = metric
        self.perplexity = perplexity
        self.max_iter_affinity = max_iter_affinity
        self.sparsity = sparsity
        super().__init__(n_components=n_components, lr=lr, optimizer=optimizer, optimizer_kwargs=optimizer_kwargs,
                         scheduler=scheduler, scheduler_kwargs=scheduler_kwargs, min_grad_norm=min_grad_norm,
                         max_iter=max_iter, init=init, init_scaling=init_scaling, device=device, backend=backend,
                         verbose=verbose, random_state=random_state,
                         early_exaggeration_coeff=early_exaggeration_coeff,
                         early_exaggeration_iter=early_exaggeration_iter, check_interval=check_interval,
                         compile=compile, distributed=distributed, **kwargs)
        self.affinity_in = EntropicAffinity(perplexity=perplexity, metric=metric, max_iter=max_iter_affinity,
                                            device=self.device, backend=backend, verbose=verbose, sparsity=sparsity,
                                            distributed=self.distributed)

    def _compute_gradient(self, Z, step):
        s, e = self.chunk_start_, self.chunk_end_
        if not hasattr(self, "_sne_rows"):
            self._sne_rows = torch.zeros((Z.shape[0], 1), dtype=torch.float32, device=Z.device)
        lam, rep = float(self.early_exaggeration_coeff_), float(self.repulsion_strength)
        ops.sne_grad(Z, s, e - s, self.affinity_in_, self.NN_indices_, lam, rep, 0, self._grad, self._sne_rows)
        if self.world_size > 1:
            # the reference has every rank compute the whole N x N term and divide by W (sne.py:172-179);
            # here each rank owns a row range and the per-row normalisers are all-gathered
            all_gather_rows(self._sne_rows, self._bounds, self.rank)
        ops.sne_grad(Z, s, e - s, self.affinity_in_, self.NN_indices_, lam, rep, 1, self._grad, self._sne_rows)

    def clear_memory(self):
        if hasattr(self, "_sne_rows"):
            del self._sne_rows
        super().clear_memory()
