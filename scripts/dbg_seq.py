"""Dev: replay the test order that precedes test_pairwise_full in one process, then dissect the mismatch."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
torch.manual_seed(42)
import oracle
from helpers import blobs
from torchdr_b200 import ops
import test_gpu_parity as T

def dissect(tag):
    X = blobs(300, 50, 3, 1)
    Xd = X.cuda()
    ref64 = torch.cdist(X.double(), X.double()) ** 2
    orc = oracle.pairwise_full(X, None, "euclidean") ** 2
    Ce = ops.pairwise_full(Xd, None, metric="euclidean").cpu() ** 2
    for name, C in (("gpu", Ce), ("oracle", orc)):
        bad = (C - ref64).abs() > 0.01
        print(tag, name, "max err", float((C - ref64).abs().max()), "n_bad", int(bad.sum()),
              "rows", bad.any(1).nonzero().flatten().tolist()[:24], "cols", bad.any(0).nonzero().flatten().tolist()[:24], flush=True)

dissect("fresh")
steps = [("golden0", lambda: T.test_knn_matches_reference_golden(ops, "knn_n300_d16_k15")),
         ("golden1", lambda: T.test_knn_matches_reference_golden(ops, "knn_n2000_d50_k90")),
         ("golden2", lambda: T.test_knn_matches_reference_golden(ops, "knn_n1500_d128_k15")),
         ("eucl", lambda: T.test_knn_euclidean_metric(ops))]
for a in [(1, 1, 1), (129, 3, 1), (257, 17, 160), (1000, 50, 7), (130, 128, 129)]:
    steps.append((f"shapes{a}", lambda a=a: T.test_knn_shapes_and_edges(ops, *a)))
steps += [("ties", lambda: T.test_knn_ties_resolve_to_lower_index(ops)), ("cross", lambda: T.test_knn_cross_and_chunk(ops))]
for name, fn in steps:
    try:
        fn()
    except Exception as e:
        print(name, "raised", type(e).__name__, str(e)[:200])
    dissect("after " + name)
