#!/usr/bin/env python
"""bench.py — UMAP iters/sec (and the fused kNN+sigma "affinity kernel" throughput) on B200.

Contract (driver): ``python bench.py --gpus N --steps K --warmup W`` (under torchrun for N>1)
prints ONE JSON line from rank 0.  ``--impl reference`` times the CPU restatement of the
reference path (oracle/, kind "port") on a bounded sample instead.

Workload (BASELINE.json configs[1]): UMAP n_neighbors=15 on 1 M x 128 synthetic clustered points
(the reference benchmark's generator, benchmarks/faiss/run_benchmark.py:127-146).  A *step* is
one UMAP optimisation iteration over all points; the graph (kNN -> sigma/rho -> symmetrise ->
edge schedule) is built once, untimed, through the same C-ABI calls.  Strong scaling: the point
set is fixed and rows are sharded across ranks; every iteration's updated rows reach the peers
through NVLink stores issued by the step kernel itself (fallback: one NCCL all-gather).
"""

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

K_NEIGHBORS = 15
N_NEG = 75          # umap.py:177: negative_sample_rate * n_neighbors
MAX_ITER = 500      # iteration budget of the schedule (the reference's UMAP benchmark uses 500)
E2E_ITERS = 500
FALLBACK_HBM_GBS = 6650.0  # /opt/skills/guides/B200_PROFILING.md


# ----------------------------------------------------------------------------- data
def clustered(n, d, device, seed=42):
    """benchmarks/faiss/run_benchmark.py:127-146: min(1000, n//100) Gaussian clusters,
    centres randn*10, points centre + randn*0.5, cluster blocks contiguous."""
    g = torch.Generator(device=device).manual_seed(seed)
    nc = max(1, min(1000, n // 100))
    centers = torch.randn(nc, d, generator=g, device=device) * 10
    per = n // nc
    lab = torch.arange(n, device=device) // per
    lab.clamp_(max=nc - 1)
    X = centers[lab] + torch.randn(n, d, generator=g, device=device) * 0.5
    return X.float().contiguous()


# ----------------------------------------------------------------------------- clocks
class ClockSampler:
    """SM clock + throttle reasons sampled DURING the timed region (NVML every 20 ms; nvidia-smi fallback)."""

    FIELDS = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
              "clocks_event_reasons.sw_power_cap")
    # nvmlClocksEventReason* bit masks
    BITS = {"sw_power_cap": 0x4, "hw_slowdown": 0x8, "sw_thermal_slowdown": 0x20, "hw_thermal_slowdown": 0x40}

    def __init__(self, index=0):
        self.index = index
        self.sm, self.max_sm, self.reasons = [], [], set()
        self._stop = threading.Event()
        self._t = None
        self._nvml = None
        try:
            import pynvml

            pynvml.nvmlInit()
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            phys = int(vis.split(",")[index]) if vis and vis.split(",")[index].isdigit() else index
            self._h = pynvml.nvmlDeviceGetHandleByIndex(phys)
            self._nvml = pynvml
            self.max_sm.append(float(pynvml.nvmlDeviceGetMaxClockInfo(self._h, pynvml.NVML_CLOCK_SM)))
        except Exception:
            self._nvml = None

    def _sample_nvml(self):
        nv = self._nvml
        self.sm.append(float(nv.nvmlDeviceGetClockInfo(self._h, nv.NVML_CLOCK_SM)))
        get = getattr(nv, "nvmlDeviceGetCurrentClocksEventReasons", None) or nv.nvmlDeviceGetCurrentClocksThrottleReasons
        mask = int(get(self._h))
        for name, bit in self.BITS.items():
            if mask & bit:
                self.reasons.add(name)

    def _sample_smi(self):
        out = subprocess.run(["nvidia-smi", f"--query-gpu={self.FIELDS}", "--format=csv,noheader,nounits",
                              "-i", str(self.index)], capture_output=True, text=True, timeout=5).stdout
        parts = [p.strip() for p in out.strip().split(",")]
        if len(parts) >= 7:
            self.sm.append(float(parts[0]))
            self.max_sm.append(float(parts[1]))
            for i, nm in enumerate(["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]):
                if parts[3 + i].lower().startswith("active"):
                    self.reasons.add(nm)

    def _run(self):
        while not self._stop.is_set():
            try:
                self._sample_nvml() if self._nvml else self._sample_smi()
            except Exception:
                pass
            self._stop.wait(0.02 if self._nvml else 0.1)

    def __enter__(self):
        self._t = threading.Thread(target=self._run, daemon=True)
        self._t.start()
        return self

    def __exit__(self, *a):
        self._stop.set()
        self._t.join(timeout=10)

    def summary(self):
        return {"sm_mhz": float(np.median(self.sm)) if self.sm else None,
                "sm_max_mhz": max(self.max_sm) if self.max_sm else None,
                "reasons": sorted(self.reasons), "samples": len(self.sm),
                "source": "nvml" if self._nvml else "nvidia-smi"}


# ----------------------------------------------------------------------------- CPU arm (oracle port)
def cpu_reference(steps, warmup, n_total, d, sample_n=20000, knn_queries=2048, verbose=False):
    """Reference path restated on CPU (oracle/, torch CPU ops like the reference's backend=None).

    Bounded sample: the full pipeline on `sample_n` points of the same generator; the loop rate is
    per-iteration over sample_n points and is scaled by sample_n / n_total (the iteration is O(N));
    the kNN stage is O(N^2 D) and is timed as `knn_queries` query rows against a database of
    sample_n rows, then scaled by (n_total/knn_queries) * (n_total/sample_n).  Extrapolated
    figures are labelled as such in `sample`."""
    import oracle

    all_cores = os.cpu_count() or 1
    torch.set_num_threads(all_cores)
    cores = torch.get_num_threads()
    X = clustered(sample_n, d, "cpu")
    t0 = time.perf_counter()
    C, I = oracle.knn_chunked(X, K_NEIGHBORS, block=4096)
    t_knn_sample = time.perf_counter() - t0
    t0 = time.perf_counter()
    oracle.knn_chunked(X, K_NEIGHBORS, block=knn_queries, q_start=0, q_end=knn_queries)
    t_knn_q = time.perf_counter() - t0
    t0 = time.perf_counter()
    P, rho, sigma = oracle.umap_affinity_rows(C, K_NEIGHBORS, max_iter=100)
    V, J = oracle.symmetrize_ell(P, I)
    per, nxt = oracle.umap_edge_schedule(V, MAX_ITER)
    t_aff = time.perf_counter() - t0
    a, b = oracle.find_ab()
    g = torch.Generator().manual_seed(0)
    Z = torch.randn(sample_n, 2, generator=g)
    Z = 1e-4 * Z / Z[:, 0].std()
    lrs = oracle.linear_lr_sequence(1.0, MAX_ITER, steps + warmup)
    me = torch.arange(sample_n)

    def one(t, Z, nxt):
        neg = torch.randint(0, sample_n - 1, (sample_n, N_NEG), generator=g)
        neg = oracle.adjust_negatives(neg, me)
        G = oracle.umap_step(Z, J, per, nxt, neg, t, a, b)
        return Z.add(G, alpha=-float(lrs[t]))

    # the loop is ~40 small ATen ops per iteration: on a many-core host all threads can be slower than a few, so
    # it is timed at min(all, 16) threads and at all threads and the faster setting is reported (both stated)
    loop_times = {}
    for nthreads in sorted({min(all_cores, 16), all_cores}):
        torch.set_num_threads(nthreads)
        Zt, nt = Z.clone(), nxt.clone()
        for t in range(warmup):
            Zt = one(t, Zt, nt)
        t0 = time.perf_counter()
        for t in range(warmup, warmup + steps):
            Zt = one(t, Zt, nt)
        loop_times[nthreads] = time.perf_counter() - t0
    loop_threads = min(loop_times, key=loop_times.get)
    t_loop = loop_times[loop_threads]
    torch.set_num_threads(all_cores)
    its_sample = steps / t_loop
    its_full = its_sample * sample_n / n_total
    knn_full_s = t_knn_q * (n_total / knn_queries) * (n_total / sample_n)
    aff_full_s = t_aff * n_total / sample_n
    e2e_full = E2E_ITERS / (knn_full_s + aff_full_s + E2E_ITERS / its_full)
    return {
        "value": its_full, "unit": "iters/s", "cores": cores, "kind": "port",
        "sample": (f"oracle (torch-CPU restatement of backend=None) on {sample_n}x{d} clustered points: loop "
                   f"{its_sample:.2f} it/s at {loop_threads} threads (timings by thread count: "
                   f"{ {k: round(v, 2) for k, v in loop_times.items()} } s) measured over {steps} iters, scaled x{sample_n}/{n_total} (O(N) per "
                   f"iteration, extrapolated); kNN {t_knn_sample:.1f}s at {sample_n} rows, {t_knn_q:.2f}s for "
                   f"{knn_queries} queries -> {knn_full_s:.0f}s extrapolated to {n_total} rows; affinity+graph "
                   f"{t_aff:.1f}s -> {aff_full_s:.0f}s extrapolated"),
        "e2e_value": e2e_full, "loop_its_sample": its_sample, "knn_full_s_extrapolated": knn_full_s,
        "ms_per_step": 1e3 * t_loop / steps * n_total / sample_n,
    }


# ----------------------------------------------------------------------------- GPU arm
def build_graph(X, rank, world, max_iter, full_sweep=True):
    """Untimed setup through the product path: fused kNN + sigma/rho, symmetrise, schedule, compact."""
    import torch.distributed as dist

    from torchdr_b200 import ops
    from torchdr_b200.distributed import all_bounds, exchange_edges

    n = X.shape[0]
    bounds = all_bounds(n, world)
    s, e = bounds[rank]
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)

    def timed_knn():
        torch.cuda.synchronize()
        ev0.record()
        out = ops.knn_umap_fused(X[s:e], X, K_NEIGHBORS, q_row0=s, want_dist=False)
        ev1.record()
        torch.cuda.synchronize()
        return out, ev0.elapsed_time(ev1)

    # default path: tile-pruned exact sweep (bit-identical to the full sweep, tests/test_gpu_parity.py); the full
    # sweep is timed as well (it is what the tensor-pipe roofline is quoted on) unless it would take minutes
    sweep = torch.zeros(2, dtype=torch.int64, device=X.device)
    knn = {}
    try:
        if full_sweep:
            ops.knn_set_prune(False)
            _, knn["full_ms"] = timed_knn()
        ops.knn_set_prune(True, None)
        timed_knn()  # warm-up: the first call pays the allocation of the (up to 10 GB) workspace
        ops.knn_set_prune(True, sweep)
        (dist_, idx, P, rho, sigma), knn["ms"] = timed_knn()
        knn["tile_pairs_swept"], knn["tile_pairs_all"] = (int(v) for v in sweep.tolist())
    finally:
        ops.knn_set_prune(True, None)
    knn_ms = knn
    ext = None
    if world > 1:
        counts, er, ec, ev = ops.symmetrize_export(P, idx, s, n, world, rank)
        ext = exchange_edges(counts, er, ec, ev)
    rowptr, col, val = ops.symmetrize_csr(P, idx, s, n, ext=ext)
    a_max = ops.max_value(val)
    if world > 1:
        dist.all_reduce(a_max, op=dist.ReduceOp.MAX)
    eps, _ = ops.umap_schedule(val, float(a_max.item()), max_iter)
    graph = ops.umap_compact(rowptr, col, eps)
    return graph, bounds, knn_ms, int(val.numel())


def gpu_arm(args):
    import torch.distributed as dist

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    from torchdr_b200 import UMAP, _lib, ops
    from torchdr_b200.distributed import all_gather_rows
    from torchdr_b200.neighbor_embedding import find_ab_params

    _lib.require_device(dev)
    n, d, K, W = args.points, args.dim, args.steps, max(args.warmup, 3)
    X = clustered(n, d, dev)
    if args.order == "shuffled":  # same points without the generator's index locality (kNN sweeps every tile)
        X = X[torch.randperm(n, generator=torch.Generator(device=dev).manual_seed(7), device=dev)].contiguous()
    sched = max(MAX_ITER, W + K + 16)  # schedule length: the timed iterations are the head of one LinearLR 1 -> 0 run
    full_sweep = args.full_sweep or n <= 2_000_000
    (rowptr, col, eps, eons), bounds, knn, nnz_sym = build_graph(X, rank, world, sched, full_sweep)
    s, e = bounds[rank]
    a, b = find_ab_params(1.0, 0.1)
    g = torch.Generator(device=dev).manual_seed(0)
    Z = torch.randn(n, 2, generator=g, device=dev)
    Za = (1e-4 * Z / Z[:, 0].std()).contiguous()
    Zb = Za.clone()
    # learning rates of the reference schedule (LinearLR 1 -> 0 over MAX_ITER), host-side scalars
    lr_all = np.asarray([1.0 * (1.0 - t / sched) for t in range(W + K)], dtype=np.float32)
    stats = torch.zeros(2, dtype=torch.int64, device=dev)
    nan_flag = torch.zeros(1, dtype=torch.int32, device=dev)

    # multi-GPU exchange: step kernel with fused NVLink peer stores when symmetric memory is available
    peer, exchange = None, ("none" if world == 1 else "nccl-allgather")
    if world > 1 and os.environ.get("TDR_NO_P2P") != "1":
        try:
            from torchdr_b200.distributed import PeerEmbedding

            peer = PeerEmbedding(Za)
            Za, Zb = peer.bufs[0], peer.bufs[1]
            exchange = "p2p-fused (tdr_umap_run_p2p_f32: step kernel with NVLink peer stores + flag barrier kernel)"
        except Exception as exc:
            if rank == 0:
                print(f"[bench] symmetric memory unavailable: {exc}", file=sys.stderr)
            peer = None

    def run(t0, count, Za, Zb, stats_t):
        if world == 1:
            res = ops.umap_run(Za, Zb, rowptr, col, eps, eons, t0, lr_all[t0:t0 + count], a, b, n_neg=N_NEG,
                               seed=1234, nan_flag=nan_flag, stats=stats_t)
            return (res, Zb if res is Za else Za)
        if peer is not None and stats_t is None and count > 0:
            cur = 0 if Za is peer.bufs[0] else 1
            cur = ops.umap_run_p2p(peer, cur, s, e - s, rowptr, col, eps, eons, t0, lr_all[t0:t0 + count], a, b,
                                   n_neg=N_NEG, seed=1234, nan_flag=nan_flag)
            return peer.bufs[cur], peer.bufs[1 - cur]
        for t in range(t0, t0 + count):  # fallback exchange: one NCCL all-gather of the updated rows per iteration
            ops.umap_step(Za, Zb, s, e - s, rowptr, col, eps, eons, t, a, b, float(lr_all[t]), neg=None,
                          n_neg=N_NEG, seed=1234, nan_flag=nan_flag, stats=stats_t)
            all_gather_rows(Zb, bounds, rank)
            Za, Zb = Zb, Za
        return Za, Zb

    Za, Zb = run(0, W, Za, Zb, None)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with ClockSampler(local_rank) as clk:
        ev0.record()
        Za, Zb = run(W, K, Za, Zb, None)
        ev1.record()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
    ms = torch.tensor([ev0.elapsed_time(ev1)], device=dev, dtype=torch.float64)
    # roofline accounting (sampled-edge / negative counters) over S extra iterations OUTSIDE the timed region
    S = min(K, 16)
    lr_all = np.concatenate([lr_all, np.full(S, lr_all[-1], dtype=np.float32)])
    Za, Zb = run(W + K, S, Za, Zb, stats)
    torch.cuda.synchronize()
    st = stats.clone().double() * (K / S)
    nnz_live = torch.tensor([float(col.numel())], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        dist.all_reduce(st, op=dist.ReduceOp.SUM)
        dist.all_reduce(nnz_live, op=dist.ReduceOp.SUM)
    assert int(nan_flag.item()) == 0, "NaN in the embedding"
    assert bool(torch.isfinite(Za).all())
    total_ms = float(ms.item())
    ms_per_step = total_ms / K
    value = 1e3 / ms_per_step

    # ---- roofline of the step kernel (algorithmic bytes) -------------
    act, negs = float(st[0]) / K, float(st[1]) / K
    nnz = float(nnz_live.item())
    alg_bytes = 16.0 * n + 8.0 * (n + world) + 4.0 * nnz + act * (4 + 4 + 8 + 4) + negs * 8.0
    peak, peak_src = FALLBACK_HBM_GBS, "fallback"
    try:
        peak = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
        peak_src = "measured"
    except Exception:
        pass
    kernel_ms = ms_per_step  # N=1: the timed region contains only the K step-kernel launches
    achieved = alg_bytes / world / (kernel_ms * 1e-3) / 1e9  # per GPU
    traffic = None
    prof = os.path.join(ROOT, "profiles", "step_kernel_traffic.json")
    if os.path.exists(prof):
        try:
            traffic = json.load(open(prof)).get("dram_bytes_per_launch")
        except Exception:
            pass
    if traffic is not None and world > 1:
        traffic = None  # the ncu capture is of the N = 1 launch
    roofline = {"bound": "hbm", "kernel": "tdr::umap_step_kernel_fast4", "achieved": achieved, "peak": peak,
                "unit": "GB/s", "frac": achieved / peak, "traffic": traffic, "peak_source": peak_src,
                "algorithmic_bytes_per_launch": alg_bytes / world,
                "note": ("per GPU; achieved = algorithmic bytes (DESIGN.md 3.5, counts taken in-kernel) / mean launch "
                         "time over the timed region; ncu shows the kernel is bound by instruction issue and the L2 "
                         "sector rate of the random z_j gathers, not by DRAM (profiles/r1_step_kernel.md); at N>1 the per-step time also contains the cross-GPU barrier")}
    aff_bytes = 4.0 * n * d / 1 + n / world * K_NEIGHBORS * 8.0 + 8.0 * n / world
    tc_peak = None
    try:
        tc_peak = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["bf16_tflops"])
    except Exception:
        pass
    affinity = {"kernel": "tdr::tc::knn_tc_kernel (fused exact kNN + sigma/rho; tcgen05 kind::f16, 3 split passes; "
                          "tile-pruned sweep)",
                "ms": knn["ms"], "algorithmic_bytes": aff_bytes, "gbs": aff_bytes / (knn["ms"] * 1e-3) / 1e9,
                "tile_pairs_swept": knn["tile_pairs_swept"], "tile_pairs_all": knn["tile_pairs_all"],
                "note": "ms = the default path: bounding-box pruned sweep, results bit-identical to the full sweep; GB/s on "
                        "the 640 MB algorithmic bytes (SURVEY 8d) is quoted because the metric asks for it. full_sweep_* = "
                        "the same kernel visiting every database tile (what the clustered generator's index locality "
                        "saves; data without locality pays it): compute-bound, the roofline that binds is the tensor pipe"}
    if "full_ms" in knn:
        tf_equiv = 2.0 * (n / world) * n * d / (knn["full_ms"] * 1e-3) / 1e12
        affinity.update({"full_sweep_ms": knn["full_ms"], "full_sweep_gbs": aff_bytes / (knn["full_ms"] * 1e-3) / 1e9,
                         "full_sweep_tflops_2nnd": tf_equiv, "full_sweep_tensor_tflops_3pass": 3.0 * tf_equiv,
                         "full_sweep_tensor_frac_of_measured_bf16_peak": (3.0 * tf_equiv / tc_peak) if tc_peak else None})

    out = None
    if rank == 0:
        out = {
            "metric": "UMAP iters/sec (1 M x 128, n_neighbors=15); affinity-kernel GB/s in `affinity_kernel`",
            "value": value, "unit": "iters/s", "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": f"UMAP n_neighbors={K_NEIGHBORS} on {n}x{d} clustered synthetic (BASELINE configs[1])",
                       "points": n, "dim": d, "row_order": args.order, "n_negatives": N_NEG, "schedule_max_iter": sched,
                       "negatives": "in-kernel Philox4x32-10", "parallelism": f"rows sharded x{world}", "exchange": exchange,
                       "l2": "per-iteration working set (CSR edge state %.0f MB) exceeds the 126 MB L2; no flush" %
                             (nnz * 12 / 1e6 / world)},
            "roofline": roofline, "affinity_kernel": affinity,
            "gpu_launches": K,  # per rank: one step-kernel launch per iteration (+ one flag-barrier kernel at N > 1)
            "clocks": clk.summary(),
            "graph": {"nnz_symmetrised": nnz_sym, "nnz_live": nnz, "sampled_edges_per_iter": act,
                      "negatives_per_iter": negs},
        }
    del X
    return out, (rank, world, dev)


def e2e_arm(args, dev):
    """fit_transform through the public estimator on HOST memory: H2D of X, kNN, sigma search, graph,
    E2E_ITERS iterations, D2H of the embedding — all inside the timed region."""
    from torchdr_b200 import UMAP

    n, d = args.points, args.dim
    Xd = clustered(n, d, dev)
    if args.order == "shuffled":
        Xd = Xd[torch.randperm(n, generator=torch.Generator(device=dev).manual_seed(7), device=dev)]
    Xh = Xd.cpu().pin_memory().numpy()
    del Xd
    torch.cuda.synchronize()
    m = UMAP(n_neighbors=K_NEIGHBORS, max_iter=E2E_ITERS, init="normal", random_state=0, process_duplicates=False)
    m.fit_transform(Xh[:20000])  # warm-up of allocator / library load
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    Z = m.fit_transform(Xh)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    assert Z.shape == (n, 2) and np.isfinite(Z).all()
    return {"value": E2E_ITERS / dt, "unit": "iters/s", "h2d_bytes_per_step": Xh.nbytes / E2E_ITERS,
            "d2h_bytes_per_step": Z.nbytes / E2E_ITERS, "seconds": dt, "iters": E2E_ITERS,
            "note": "UMAP(n_neighbors=15, max_iter=500, init='normal').fit_transform(numpy X): "
                    "iters / wall time incl. H2D, exact kNN, sigma search, symmetrise, loop, D2H"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=2000)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--points", type=int, default=1_000_000)
    ap.add_argument("--dim", type=int, default=128)
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--full-sweep", action="store_true", help="also time the unpruned kNN sweep above 2 M points")
    ap.add_argument("--order", default="generator", choices=["generator", "shuffled"],
                    help="row order of the synthetic points: the reference generator's (clusters contiguous) or shuffled")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    # NCCL prints its version banner to STDOUT when NCCL_DEBUG=VERSION; the contract is ONE JSON line there
    if os.environ.get("NCCL_DEBUG", "").upper() == "VERSION":
        os.environ["NCCL_DEBUG"] = "WARN"
    # ... and native libraries may still write to fd 1: park the real stdout and point fd 1 at stderr until the end
    sys.stdout.flush()
    real_stdout = os.dup(1)
    os.dup2(2, 1)

    def emit(obj):
        sys.stdout.flush()
        os.write(real_stdout, (json.dumps(obj) + "\n").encode())

    if args.impl == "reference":
        if rank != 0:
            return
        steps = min(args.steps, 20)
        r = cpu_reference(steps, min(args.warmup, 3), args.points, args.dim)
        line = {
            "impl": "reference",
            "metric": "UMAP iters/sec (1 M x 128, n_neighbors=15); affinity-kernel GB/s in `affinity_kernel`",
            "value": r["value"], "unit": "iters/s", "n_gpus": args.gpus, "steps": steps, "warmup": min(args.warmup, 3),
            "ms_per_step": r["ms_per_step"], "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": f"UMAP n_neighbors={K_NEIGHBORS} on {args.points}x{args.dim} clustered synthetic "
                                   "(BASELINE configs[1])", "points": args.points, "dim": args.dim},
            "cpu_baseline": {"value": r["value"], "unit": "iters/s", "cores": r["cores"], "kind": r["kind"],
                             "sample": r["sample"]},
            "e2e": {"value": r["e2e_value"], "unit": "iters/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        }
        emit(line)
        return

    out, (rank, world, dev) = gpu_arm(args)
    e2e = None
    if not args.no_e2e:
        # every rank calls the estimator (it shards rows and runs its collectives); rank 0 reports its wall time,
        # which contains every collective of the fit
        e2e = e2e_arm(args, dev)
    if rank == 0:
        if e2e is not None:
            out["e2e"] = e2e
        if world == 1 and not args.no_cpu:
            r = cpu_reference(10, 2, args.points, args.dim)
            out["cpu_baseline"] = {"value": r["value"], "unit": "iters/s", "cores": r["cores"], "kind": r["kind"],
                                   "sample": r["sample"], "e2e_value": r["e2e_value"]}
        emit(out)
    if world > 1:
        import torch.distributed as dist

        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
