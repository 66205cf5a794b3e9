"""Dev: stage timings of UMAP.fit_transform on host input (TDR_TIMING=1 synchronises after every stage)."""
import os, sys, time
os.environ["TDR_TIMING"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from bench import clustered
from torchdr_b200 import UMAP
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
order = sys.argv[2] if len(sys.argv) > 2 else "generator"
Xd = clustered(n, 128, "cuda")
if order == "shuffled":
    Xd = Xd[torch.randperm(n, device="cuda", generator=torch.Generator(device="cuda").manual_seed(7))]
Xh = Xd.cpu().pin_memory().numpy()
del Xd
m = UMAP(n_neighbors=15, max_iter=500, init="normal", random_state=0, process_duplicates=False)
m.fit_transform(Xh[:20000])
for rep in range(2):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    Z = m.fit_transform(Xh)
    torch.cuda.synchronize(); dt = time.perf_counter() - t0
    t = m.timings_
    t = dict(t, **{"(of which knn+sigma": getattr(m.affinity_in, "timings_", {}).get("knn+sigma", float("nan"))})
    print(f"n={n} order={order} total {dt*1e3:.1f} ms (with per-stage syncs): " + ", ".join(f"{k} {v*1e3:.1f}" for k, v in t.items()), flush=True)
