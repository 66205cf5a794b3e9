"""Dev: one fused tensor-core kNN launch at n=100k for ncu."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import torch
from torchdr_b200 import ops, _lib
from helpers import clustered
n = int(sys.argv[1]) if len(sys.argv) > 1 else 100_000
path = int(sys.argv[2]) if len(sys.argv) > 2 else 2
_lib.load().tdr_knn_set_path(path)
X = clustered(n, 128).to("cuda:0")
for _ in range(2):
    out = ops.knn_umap_fused(X, X, 15)
torch.cuda.synchronize()
print("done")
