"""Multi-GPU check (run under torchrun, NCCL): the row-sharded path against the single-GPU path.

  python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 scripts/dist_check.py
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import numpy as np
import torch
import torch.distributed as dist

rank = int(os.environ["RANK"])
world = int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
dev = torch.device("cuda", int(os.environ["LOCAL_RANK"]))
dist.init_process_group("nccl", device_id=dev)

import torchdr_b200 as tb
from torchdr_b200 import ops
from torchdr_b200.distributed import all_bounds, all_gather_rows, exchange_edges
from helpers import blobs

n, d, k = 20011, 48, 15
X = blobs(n, d, 12, 5).to(dev)
bounds = all_bounds(n, world)
s, e = bounds[rank]
ok = True


def report(name, cond):
    global ok
    ok = ok and bool(cond)
    if rank == 0:
        print(f"[dist_check] {name}: {'ok' if cond else 'FAIL'}", flush=True)


# ---- graph: chunk kNN + sigma, edge exchange over NCCL, local symmetrise == rows of the single-GPU graph
dist_f, idx_f, P_f, rho_f, sig_f = ops.knn_umap_fused(X, X, k)
rp_f, col_f, val_f = ops.symmetrize_csr(P_f, idx_f, 0, n)
dist_c, idx_c, P_c, rho_c, sig_c = ops.knn_umap_fused(X[s:e], X, k, q_row0=s)
report("chunk kNN == full kNN rows", torch.equal(idx_c, idx_f[s:e]) and torch.equal(P_c, P_f[s:e]))
counts, er, ec, ev = ops.symmetrize_export(P_c, idx_c, s, n, world, rank)
ext = exchange_edges(counts, er, ec, ev)
rp_c, col_c, val_c = ops.symmetrize_csr(P_c, idx_c, s, n, ext=ext)
a, b = int(rp_f[s]), int(rp_f[e])
report("distributed symmetrise == single-GPU rows",
       torch.equal(rp_c, rp_f[s:e + 1] - rp_f[s]) and torch.equal(col_c, col_f[a:b]) and torch.equal(val_c, val_f[a:b]))

# ---- loop: 8 sharded iterations with all-gather == 8 single-GPU iterations
a_max = ops.max_value(val_f)
eps_f, _ = ops.umap_schedule(val_f, float(a_max.item()), 100)
g_full = ops.umap_compact(rp_f, col_f, eps_f)
a_max_c = ops.max_value(val_c)
dist.all_reduce(a_max_c, op=dist.ReduceOp.MAX)
report("A_max all-reduce", float(a_max_c.item()) == float(a_max.item()))
eps_c, _ = ops.umap_schedule(val_c, float(a_max_c.item()), 100)
g_loc = ops.umap_compact(rp_c, col_c, eps_c)
gen = torch.Generator(device=dev).manual_seed(3)
Z0 = (torch.randn(n, 2, generator=gen, device=dev) * 1e-4).contiguous()
dist.broadcast(Z0, src=0)
from torchdr_b200.neighbor_embedding import find_ab_params

pa, pb = find_ab_params(1.0, 0.1)
lrs = np.linspace(1.0, 0.9, 8).astype(np.float32)
Za, Zb = Z0.clone(), torch.empty_like(Z0)
res = ops.umap_run(Za, Zb, *g_full[:3], g_full[3].clone(), 0, lrs, pa, pb, seed=11)
Zc, Zd = Z0.clone(), Z0.clone()
eons = g_loc[3].clone()
for t in range(8):
    ops.umap_step(Zc, Zd, s, e - s, g_loc[0], g_loc[1], g_loc[2], eons, t, pa, pb, float(lrs[t]), neg=None, seed=11)
    all_gather_rows(Zd, bounds, rank)
    Zc, Zd = Zd, Zc
err = float((Zc - res).norm() / res.norm())
report(f"8 sharded iterations vs single GPU (rel {err:.2e})", err < 1e-3)

# ---- fused step + exchange over NVLink peer stores (symmetric memory) == NCCL all-gather path
try:
    from torchdr_b200.distributed import PeerEmbedding

    peer = PeerEmbedding(Z0.numel(), Z0.dtype, Z0.device)
    peer.load(Z0)
    Zp, Zq = peer.bufs[0], peer.bufs[1]
    eons2 = g_loc[3].clone()
    cur = 0
    for t in range(8):
        ops.umap_step_p2p(Zp, Zq, s, e - s, g_loc[0], g_loc[1], g_loc[2], eons2, t, pa, pb, float(lrs[t]),
                          peer.peer_ptrs(1 - cur), seed=11)
        peer.barrier(1 - cur)
        Zp, Zq = Zq, Zp
        cur = 1 - cur
    report("p2p-fused step: 8 iterations bit-identical to the all-gather path", torch.equal(Zp, Zc))
    # native multi-step loop (flag barrier kernel instead of the host-driven barrier)
    peer2 = PeerEmbedding(Z0.numel(), Z0.dtype, Z0.device)
    peer2.load(Z0)
    c2 = ops.umap_run_p2p(peer2, 0, s, e - s, g_loc[0], g_loc[1], g_loc[2], g_loc[3].clone(), 0, lrs, pa, pb, seed=11)
    torch.cuda.synchronize()
    report("native p2p loop (tdr_umap_run_p2p_f32): 8 iterations bit-identical", torch.equal(peer2.bufs[c2], Zc))
    import time
    e3 = g_loc[3].clone()
    torch.cuda.synchronize(); dist.barrier(); t0 = time.perf_counter()
    c2 = ops.umap_run_p2p(peer2, c2, s, e - s, g_loc[0], g_loc[1], g_loc[2], e3, 8, [0.5] * 200, pa, pb, seed=11)
    torch.cuda.synchronize(); t_nat = (time.perf_counter() - t0) / 200
    if rank == 0:
        print(f"[dist_check] native p2p loop per-iteration wall time at n={n}: {t_nat*1e6:.1f} us", flush=True)
    torch.cuda.synchronize(); dist.barrier(); t0 = time.perf_counter()
    for t in range(200):
        ops.umap_step_p2p(Zp, Zq, s, e - s, g_loc[0], g_loc[1], g_loc[2], eons2, 8 + t, pa, pb, 0.5, peer.peer_ptrs(1 - cur), seed=11)
        peer.barrier(1 - cur); Zp, Zq = Zq, Zp; cur = 1 - cur
    torch.cuda.synchronize(); t_p2p = (time.perf_counter() - t0) / 200
    Zc2, Zd2 = Zc.clone(), Zc.clone()
    torch.cuda.synchronize(); dist.barrier(); t0 = time.perf_counter()
    for t in range(200):
        ops.umap_step(Zc2, Zd2, s, e - s, g_loc[0], g_loc[1], g_loc[2], eons, 8 + t, pa, pb, 0.5, neg=None, seed=11)
        all_gather_rows(Zd2, bounds, rank); Zc2, Zd2 = Zd2, Zc2
    torch.cuda.synchronize(); t_ag = (time.perf_counter() - t0) / 200
    if rank == 0:
        print(f"[dist_check] per-iteration wall time at n={n}: p2p-fused {t_p2p*1e6:.1f} us, nccl all-gather {t_ag*1e6:.1f} us", flush=True)
except Exception as exc:
    report(f"p2p-fused path unavailable: {type(exc).__name__}: {exc}", False)

# ---- estimators under torchrun
Xh = X.cpu().numpy()
for cls, kw in ((tb.UMAP, dict(n_neighbors=15, max_iter=60)), (tb.LargeVis, dict(perplexity=10, max_iter=40)),
                (tb.TSNE, dict(perplexity=10, max_iter=30)), (tb.InfoTSNE, dict(perplexity=10, max_iter=30, n_negatives=64)),
                (tb.SNE, dict(perplexity=10, max_iter=30, lr=30.0))):
    m = cls(init="normal", random_state=0, process_duplicates=False, **kw)
    Z = m.fit_transform(Xh)
    fin = bool(np.isfinite(Z).all()) and Z.shape == (n, 2)
    Zt = torch.from_numpy(Z).to(dev)
    Z0r = Zt.clone()
    dist.broadcast(Z0r, src=0)
    same = float((Zt - Z0r).abs().max())
    report(f"{cls.__name__} fit under torchrun x{world}: finite={fin}, max |rank0 - rank{rank}| = {same:.1e}", fin and same == 0.0)

# rows WITHOUT index locality: the fit runs in rank 0's Voronoi-tree order (broadcast), result in input order on all ranks
Xs = blobs(30011, 48, 40, 7)
Xs = Xs[torch.randperm(30011, generator=torch.Generator().manual_seed(1))].contiguous()
from torchdr_b200 import reorder as _ro
seen = {}
_orig = tb.UMAPAffinity.compute_csr
def _spy(self, X):
    seen["order"] = self.knn_order
    return _orig(self, X)
tb.UMAPAffinity.compute_csr = _spy
m = tb.UMAP(n_neighbors=15, max_iter=40, init="normal", random_state=0, process_duplicates=False)
Zs = m.fit_transform(Xs.numpy())
tb.UMAPAffinity.compute_csr = _orig
Zt = torch.from_numpy(Zs).to(dev)
Z0r = Zt.clone()
dist.broadcast(Z0r, src=0)
report(f"UMAP on shuffled rows x{world}: order={seen.get('order')}, locality {_ro.index_locality(Xs.to(dev)):.2f}, identical on all ranks",
       seen.get("order") == "presorted" and bool(np.isfinite(Zs).all()) and float((Zt - Z0r).abs().max()) == 0.0)
# LargeVis: row-local form (default) vs the scatter + all-reduce form, 6 iterations from the same initialisation
g1 = torch.Generator().manual_seed(5)
Zi2 = torch.randn(n, 2, generator=g1)
Za_ = tb.LargeVis(perplexity=10, max_iter=6, init=Zi2, random_state=0, process_duplicates=False, knn_order="input").fit_transform(Xh)
Zb_ = tb.LargeVis(perplexity=10, max_iter=6, init=Zi2, random_state=0, process_duplicates=False, knn_order="input", row_local=False).fit_transform(Xh)
Zc_ = tb.LargeVis(perplexity=10, max_iter=6, init=Zi2, random_state=0, process_duplicates=False, knn_order="input", distributed=False).fit_transform(Xh)
rel_ab = float(np.linalg.norm(Za_ - Zb_) / np.linalg.norm(Zb_)); rel_ac = float(np.linalg.norm(Za_ - Zc_) / np.linalg.norm(Zc_))
report(f"LargeVis x{world}: row-local vs scatter+all-reduce (rel {rel_ab:.2e}), vs single-GPU row-local (rel {rel_ac:.2e})", rel_ab < 1e-4 and rel_ac < 1e-4)

# row-sharded dense SNE / sampled InfoTSNE vs the same fit on one GPU (short run, injected init)
g0 = torch.Generator().manual_seed(3)
Zi = torch.randn(n, 2, generator=g0)
for cls, kw in ((tb.SNE, dict(perplexity=10, max_iter=5, lr=30.0)), (tb.InfoTSNE, dict(perplexity=10, max_iter=5, n_negatives=64))):
    Zd = cls(init=Zi, random_state=0, process_duplicates=False, **kw).fit_transform(Xh)
    Zs = cls(init=Zi, random_state=0, process_duplicates=False, distributed=False, **kw).fit_transform(Xh)
    rel = float(np.linalg.norm(Zd - Zs) / np.linalg.norm(Zs))
    report(f"{cls.__name__}: {world}-GPU fit vs single-GPU fit after 5 iterations (rel {rel:.2e})", rel < 1e-3)

t = torch.tensor([1.0 if ok else 0.0], device=dev)
dist.all_reduce(t, op=dist.ReduceOp.MIN)
if rank == 0:
    print("[dist_check] ALL OK" if float(t.item()) == 1.0 else "[dist_check] FAILURES", flush=True)
dist.barrier()
dist.destroy_process_group()
