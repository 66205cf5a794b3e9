"""ctypes binding of libtdrb200.so (the C ABI declared in include/tdrb200.h).

PyTorch is used for device memory and streams only: every call passes raw device
pointers (``tensor.data_ptr()``) and ``torch.cuda.current_stream()``.  There is no
CPU fallback — if the shared library is missing or the device is not a B200-class
GPU the call raises.
"""

import ctypes
import os
from ctypes import c_char_p, c_double, c_float, c_int, c_int64, c_size_t, c_uint64, c_void_p

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "libtdrb200.so")

TDR_OK = 0
TDR_E_INVALID = -1
TDR_MAX_K = 160
TDR_RUN_SYNC_WORDS = 8
TDR_RUN_STATUS_WORD = 4
TDR_TREE_SPARE_NODES = 2
METRIC_IDS = {"sqeuclidean": 0, "euclidean": 1}
SYM_MODES = {"sum_minus_prod": 0, "sum": 1}
KNN_PATHS = {"auto": 0, "simt": 1, "tc": 2}
KNN_PRUNE = {"default": -1, "off": 0, "on": 1, "certified": 2}

P = c_void_p  # every device pointer

# name -> (restype, argtypes); mirrors include/tdrb200.h one to one
SIGNATURES = {
    "tdr_abi_version": (c_int, []),
    "tdr_last_error": (c_char_p, []),
    "tdr_device_info": (c_int, [ctypes.POINTER(c_int)] * 3),
    "tdr_knn_workspace_bytes": (c_size_t, [c_int64, c_int64, c_int, c_int]),
    "tdr_knn_f32": (c_int, [P, c_int64, c_int64, P, c_int64, c_int, c_int, c_int, c_int, P, P, c_int, c_int, P, P, P,
                            c_size_t, P]),
    "tdr_pairwise_full_f32": (c_int, [P, c_int64, P, c_int64, c_int, c_int, c_int, P, c_int, P, c_size_t, P]),
    "tdr_tree_assign_f32": (c_int, [P, c_int, P, P, c_int64, P, P, c_int, P, P]),
    "tdr_tree_accumulate_f32": (c_int, [P, c_int, P, P, P, c_int64, c_int64, c_int, P, P, P]),
    "tdr_umap_affinity_f32": (c_int, [P, c_int64, c_int, c_int, P, P, P, P]),
    "tdr_entropic_affinity_f32": (c_int, [P, c_int64, c_int, c_float, c_float, c_int, c_float, c_float, c_float,
                                          c_float, c_int, P, P, P, P]),
    "tdr_indexed_dist_f32": (c_int, [P, P, c_int64, P, c_int64, c_int, P, c_int, c_int, c_int, P, P]),
    "tdr_entropic_dense_f32": (c_int, [P, c_int64, c_int64, c_float, c_float, c_int, c_float, c_float, c_float,
                                       c_float, c_int, P, P, P, P]),
    "tdr_knn_umap_fused_f32": (c_int, [P, c_int64, c_int64, P, c_int64, c_int, c_int, c_int, c_int, P, P, P, P, P,
                                       c_int, c_int, P, P, P, c_size_t, P]),
    "tdr_symmetrize_workspace_bytes": (c_size_t, [c_int64, c_int, c_int64]),
    "tdr_symmetrize_csr_f32": (c_int, [P, P, c_int64, c_int, c_int64, c_int64, P, P, P, c_int64, c_int, c_int, P, P, P,
                                       P, P, c_size_t, P]),
    "tdr_symmetrize_export_f32": (c_int, [P, P, c_int64, c_int, c_int64, c_int64, c_int, c_int, P, P, P, P, P]),
    "tdr_csr_to_ell_f32": (c_int, [P, P, P, c_int64, c_int64, c_float, P, P, P]),
    "tdr_max_f32": (c_int, [P, c_int64, P, P]),
    "tdr_umap_schedule_f32": (c_int, [P, c_int64, c_float, c_int, P, P, P]),
    "tdr_compact_workspace_bytes": (c_size_t, [c_int64, c_int64]),
    "tdr_umap_compact_f32": (c_int, [P, P, P, c_int64, c_int64, P, P, P, P, P, P, c_size_t, P]),
    "tdr_umap_step_f32": (c_int, [P, P, c_int64, c_int64, c_int64, P, P, P, P, P, c_int, c_int, c_uint64, c_int64,
                                  c_double, c_double, c_float, c_float, c_float, c_int, P, P, P, P, P]),
    "tdr_umap_run_f32": (c_int, [P, P, c_int64, P, P, P, P, c_int, c_int, c_uint64, c_int64, c_int,
                                 ctypes.POINTER(c_float), c_double, c_double, c_float, c_float, c_int, P, P, P, P,
                                 ctypes.c_uint32, P]),
    "tdr_umap_step_p2p_f32": (c_int, [P, P, c_int64, c_int64, c_int64, P, P, P, P, c_int, c_int, c_uint64, c_int64,
                                      c_double, c_double, c_float, c_float, c_float, P, P,
                                      ctypes.POINTER(c_uint64), c_int, P]),
    "tdr_umap_run_p2p_f32": (c_int, [P, P, c_int64, c_int64, c_int64, P, P, P, P, c_int, c_int, c_uint64, c_int64, c_int,
                                     ctypes.POINTER(c_float), c_double, c_double, c_float, c_float, P, P, P, P,
                                     ctypes.POINTER(c_uint64), ctypes.POINTER(c_uint64), P, ctypes.POINTER(c_uint64),
                                     c_int, c_int, c_int, ctypes.c_uint32, c_double, P]),
    "tdr_largevis_grad_f32": (c_int, [P, c_int64, c_int64, c_int64, P, P, c_int, P, c_int, c_uint64, c_int64,
                                      c_float, c_float, P, P]),
    "tdr_largevis_step_f32": (c_int, [P, P, c_int64, c_int64, c_int64, P, P, P, c_int, c_uint64, c_int64, c_float,
                                      c_float, P, P, c_float, c_float, c_int, P, P, ctypes.POINTER(c_uint64), c_int, P]),
    "tdr_tsne_workspace_bytes": (c_size_t, [c_int64]),
    "tdr_tsne_grad_f32": (c_int, [P, c_int64, c_int64, c_int64, P, P, c_int, c_float, c_float, c_int, P, P, c_size_t,
                                  P]),
    "tdr_infotsne_grad_f32": (c_int, [P, c_int64, c_int64, c_int64, P, P, c_int, P, c_int, c_uint64, c_int64,
                                      c_float, c_float, P, P]),
    "tdr_sne_grad_f32": (c_int, [P, c_int64, c_int64, c_int64, P, P, c_int, c_float, c_float, c_int, P, P, P]),
    "tdr_sgd_momentum_f32": (c_int, [P, P, P, c_int64, c_float, c_float, c_int, P, P, P]),
}

_lib = None


class B200EngineError(RuntimeError):
    """Raised when the native engine is unavailable or a call fails."""


def load():
    """Load the shared library (no device needed) and bind every declared symbol."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise B200EngineError(
            f"[TorchDR-B200] native library not found at {LIB_PATH}; run `python -c \"import __graft_entry__ as g; "
            "g.build()\"` (or ./build.sh).  There is no CPU fallback."
        )
    lib = ctypes.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if the .so does not export it
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def last_error():
    return load().tdr_last_error().decode("utf-8", "replace")


def check(rc, what=""):
    if rc == TDR_OK:
        return
    msg = last_error()
    if rc == TDR_E_INVALID:
        raise ValueError(msg if msg.startswith("[TorchDR]") else f"[TorchDR] ERROR : {msg}")
    raise B200EngineError(f"[TorchDR-B200] {what} failed (code {rc}): {msg}")


_device_ok = {}


def require_device(device):
    """The engine runs on sm_100 only; anything else is an error, not a fallback."""
    if not torch.cuda.is_available():
        raise B200EngineError("[TorchDR-B200] no CUDA device visible: the engine has no CPU path.")
    device = torch.device(device)
    if device.type != "cuda":
        raise B200EngineError(f"[TorchDR-B200] tensors must live on a CUDA device, got {device}.")
    idx = device.index if device.index is not None else torch.cuda.current_device()
    if idx not in _device_ok:
        with torch.cuda.device(idx):
            sm, major, minor = c_int(), c_int(), c_int()
            check(load().tdr_device_info(ctypes.byref(sm), ctypes.byref(major), ctypes.byref(minor)), "device probe")
            _device_ok[idx] = (sm.value, major.value, minor.value)
    return _device_ok[idx]


def ptr(t):
    return None if t is None else c_void_p(t.data_ptr())


def stream():
    return c_void_p(torch.cuda.current_stream().cuda_stream)


def workspace(nbytes, device):
    """256-byte aligned scratch buffer (torch's caching allocator aligns to 512)."""
    return torch.empty(max(int(nbytes), 256), dtype=torch.uint8, device=device)
