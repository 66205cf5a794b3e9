"""Dev: print every rel_fro value the loop-parity tests compare, over several runs (atomics order varies)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
torch.manual_seed(42)
from torchdr_b200 import ops
import helpers
import test_gpu_parity as T
vals = []
orig = helpers.rel_fro
def spy(a, b):
    v = orig(a, b); vals.append(v); return v
T.rel_fro = spy
for name in ("test_largevis_gradient_and_steps", "test_tsne_gradient_and_steps", "test_infotsne_gradient_and_steps",
             "test_sne_gradient_and_steps"):
    for rep in range(8):
        vals.clear()
        try:
            getattr(T, name)(ops)
            st = "ok"
        except TypeError:
            try:
                getattr(T, name)(); st = "ok"
            except AssertionError as e:
                st = "FAIL " + str(e)[:60]
        except AssertionError as e:
            st = "FAIL " + str(e)[:60]
        print(name, rep, st, " ".join(f"{v:.1e}" for v in vals), flush=True)
