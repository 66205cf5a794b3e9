"""Dev: structure of the pairwise_full mismatch seen on the GPU box."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
import oracle
from helpers import blobs
from torchdr_b200 import ops
X = blobs(300, 50, 3, 1)
Xd = X.cuda()
ref64 = torch.cdist(X.double(), X.double()) ** 2
orc = oracle.pairwise_full(X, None, "euclidean") ** 2
orc_sq = oracle.pairwise_full(X, None, "sqeuclidean")
print("cpu threads", torch.get_num_threads(), "oracle eucl^2 vs fp64 max", float((orc - ref64).abs().max()),
      "oracle sq vs fp64 max", float((orc_sq - ref64).abs().max()))
for it in range(3):
    Ce = ops.pairwise_full(Xd, None, metric="euclidean").cpu() ** 2
    Cs = ops.pairwise_full(Xd, None, metric="sqeuclidean").cpu()
    for name, C in (("eucl^2", Ce), ("sq", Cs)):
        err = (C - ref64).abs()
        bad = err > 0.01
        rows = bad.any(1).nonzero().flatten().tolist()
        cols = bad.any(0).nonzero().flatten().tolist()
        print(it, name, "gpu vs fp64 max", float(err.max()), "n_bad", int(bad.sum()), "rows", rows[:40], "cols", cols[:40])
    errO = (orc - ref64).abs() > 0.01
    print(it, "oracle bad", int(errO.sum()), "rows", errO.any(1).nonzero().flatten().tolist()[:40])
