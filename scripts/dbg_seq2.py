"""Dev: hammer test_pairwise_full's call sequence and dissect any mismatch."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
torch.manual_seed(42)
import oracle
from helpers import blobs, golden, t
from torchdr_b200 import ops
g = golden("pairwise_full_n64")
X64, Y64 = t(g["X"]), t(g["Y"])
X = blobs(300, 50, 3, 1)
ref64 = torch.cdist(X.double(), X.double()) ** 2
nbad_runs = 0
for it in range(300):
    C = ops.pairwise_full(X64.to("cuda:0"), None, exclude_diag=True).cpu()
    Cxy = ops.pairwise_full(X64.to("cuda:0"), Y64.to("cuda:0")).cpu()
    Ce = ops.pairwise_full(X.to("cuda:0"), None, metric="euclidean").cpu()
    bad = (Ce ** 2 - ref64).abs() > 0.01
    if bool(bad.any()):
        nbad_runs += 1
        if nbad_runs <= 5:
            print(it, "n_bad", int(bad.sum()), "rows", bad.any(1).nonzero().flatten().tolist()[:40], "cols",
                  bad.any(0).nonzero().flatten().tolist()[:40], flush=True)
            r = bad.any(1).nonzero().flatten()[0].item()
            print("   row", r, "gpu", (Ce[r, :6] ** 2).tolist(), "ref", ref64[r, :6].tolist())
print("bad runs:", nbad_runs, "of 300")
