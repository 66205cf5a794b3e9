"""CPU checks of the boundary: the C-ABI library loads and exports every declared symbol, the
product never imports the oracle, and the product fails loudly without a GPU."""

import ctypes
import os
import re
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "tdrb200.h")


def _declared():
    src = open(HEADER).read()
    return sorted(set(re.findall(r"TDR_API\s+[a-z_ \*]+?\b(tdr_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    from torchdr_b200 import _lib

    names = _declared()
    assert len(names) >= 20
    lib = ctypes.CDLL(_lib.LIB_PATH)
    missing = [n for n in names if not hasattr(lib, n)]
    assert not missing, missing
    assert sorted(_lib.SIGNATURES) == names  # the ctypes table binds exactly the header
    assert _lib.load().tdr_abi_version() == 2


def test_shared_library_is_sm100a_native():
    so = os.path.join(ROOT, "torchdr_b200", "lib", "libtdrb200.so")
    out = subprocess.run(["cuobjdump", "-lelf", so], capture_output=True, text=True)
    if out.returncode != 0:
        pytest.skip("cuobjdump unavailable")
    assert "sm_100a" in out.stdout


def test_no_torch_types_in_header():
    src = open(HEADER).read()
    code = re.sub(r"/\*.*?\*/", "", src, flags=re.S)  # comments may mention PyTorch; declarations may not
    assert "torch" not in code.lower() and "at::" not in code and "#include <cuda" not in code
    assert set(re.findall(r"#include <([a-z_.]+)>", code)) <= {"stddef.h", "stdint.h"}


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "torchdr_b200")
    for base, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(base, f)).read()
                assert not re.search(r"^\s*(import|from)\s+oracle", txt, re.M), f
                assert "oracle/" not in txt, f


def test_fails_loudly_without_gpu():
    import torch

    if torch.cuda.is_available():
        pytest.skip("GPU present")
    import torchdr_b200 as tb
    from torchdr_b200._lib import B200EngineError

    X = np.random.randn(40, 5).astype(np.float32)
    for call in (lambda: tb.pairwise_distances(X, k=3),
                 lambda: tb.UMAPAffinity(n_neighbors=5)(X),
                 lambda: tb.UMAP(n_neighbors=5, max_iter=2).fit_transform(X),
                 lambda: tb.TSNE(perplexity=5, max_iter=2).fit_transform(X)):
        with pytest.raises(B200EngineError, match="no CPU path"):
            call()


def test_reference_paths_not_read_at_runtime():
    for rel in ("bench.py", "__graft_entry__.py"):
        p = os.path.join(ROOT, rel)
        if os.path.exists(p):
            assert "/root/reference" not in open(p).read(), rel


def test_every_entry_point_rejects_null_arguments_before_touching_cuda():
    """Error behaviour of the boundary, checkable without a GPU: every compute entry point validates its arguments
    first and answers TDR_E_INVALID with a message naming itself — never a crash, never a CUDA call."""
    from torchdr_b200 import _lib

    lib = _lib.load()
    checked = 0
    for name, (res, args) in sorted(_lib.SIGNATURES.items()):
        if res is not ctypes.c_int or name in ("tdr_abi_version", "tdr_device_info"):
            continue
        zeros = []
        for t in args:
            try:
                zeros.append(None if t in (ctypes.c_void_p, ctypes.c_char_p) else t(0))
            except TypeError:
                zeros.append(None)  # pointer-typed argument
        assert getattr(lib, name)(*zeros) == _lib.TDR_E_INVALID, name
        assert _lib.last_error().startswith(("tdr_", "knn:")), (name, _lib.last_error())
        checked += 1
    assert checked >= 24


def test_knn_argument_errors_carry_the_reference_wording():
    """`torch.py:63-64`: an unsupported metric is a ValueError with the reference's text; k beyond the available neighbours
    and k outside the kernel's range are refused with the numbers in the message."""
    from torchdr_b200 import _lib

    lib = _lib.load()
    p = ctypes.c_void_p(256)  # never dereferenced: validation comes first

    def knn(ndb, d, k, exclude_self, metric):
        return lib.tdr_knn_f32(p, 10, 0, p, ndb, d, k, exclude_self, metric, p, p, 0, -1, None, None, None, 0, None)

    assert knn(10, 4, 3, 1, 7) == _lib.TDR_E_INVALID
    assert _lib.last_error() == "[TorchDR] ERROR : metric id 7 is not supported."
    assert knn(10, 4, 10, 1, 0) == _lib.TDR_E_INVALID and "exceeds the 9 available" in _lib.last_error()
    assert knn(10, 4, 10, 0, 0) != _lib.TDR_E_INVALID or "exceeds" not in _lib.last_error()  # k = ndb is fine without self exclusion
    assert knn(1000, 4, 161, 1, 0) == _lib.TDR_E_INVALID and "outside [1,160]" in _lib.last_error()
    with pytest.raises(ValueError, match=r"\[TorchDR\] ERROR"):
        _lib.check(knn(10, 4, 3, 1, 7), "tdr_knn_f32")
