"""Dev: accuracy of the tensor-core kNN (err / scale vs fp64) for the current TDR_TC_DEBUG variant."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import torch
from torchdr_b200 import ops, _lib
from helpers import blobs, clustered
import oracle

for (n, d, k, gen) in [(1500, 128, 15, "blobs"), (4096, 96, 15, "blobs"), (20000, 128, 15, "clustered")]:
    X = blobs(n, d, 6, n + d) if gen == "blobs" else clustered(n, d)
    C, I = ops.knn(X.cuda(), X.cuda(), k, path="tc")
    idx64, d64, ok, _ = oracle.knn_ambiguity(X, k)
    scale = float((X ** 2).sum(1).max()) * 2
    err = float((C.cpu().double() - d64).abs().max()) / scale
    bias = float((C.cpu().double() - d64).mean()) / scale
    print(f"debug={os.environ.get('TDR_TC_DEBUG','0')} n={n} d={d}: max err/scale {err:.2e} mean signed {bias:+.2e} decided-match {bool(torch.equal(I.cpu().long()[ok], idx64[ok]))}", flush=True)
