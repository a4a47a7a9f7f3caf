"""ctypes binding of libdewi_b200.so (the C ABI declared in include/dewi_b200.h).

There is no CPU fallback: if the shared library is missing, or no B200 is visible, the entry
points raise (`NativeUnavailable`, an ImportError, mirroring how the reference's optional backends
fail in their constructors, src/dewi/backends.py:171-172,249-250).
"""

from __future__ import annotations

import ctypes
import os
from ctypes import POINTER, c_char_p, c_double, c_float, c_int, c_int32, c_int64, c_size_t, c_uint32, c_uint64, c_void_p
from pathlib import Path

LIB_NAME = "libdewi_b200.so"
LIB_PATH = Path(__file__).resolve().parent / LIB_NAME

SPACE = {"cosine": 0, "l2": 1}
DTYPE = {"fp32": 0, "float32": 0, "bf16": 1, "bfloat16": 1}
FLAG_QUERY_NORMALIZED = 1 << 0
FLAG_FORCE_SIMT = 1 << 1
FLAG_FORCE_TC = 1 << 2
FLAG_HOST_IO = 1 << 3
FLAG_PRECISE_QUERY = 1 << 4
FLAG_SCOPE_FULL = 1 << 5
FLAG_NO_PAIR = 1 << 6
FLAG_NO_SEED = 1 << 7
FLAG_NO_M64 = 1 << 8
FLAG_NO_CERT = 1 << 9
FLAG_FORCE_CERT = 1 << 10
FLAG_NO_ROWS_ON_M = 1 << 11
JOIN_BF16 = 1 << 0
JOIN_FORCE_SIMT = 1 << 1
JOIN_FORCE_TC = 1 << 2
JOIN_NO_SYMMETRY = 1 << 3


class NativeUnavailable(ImportError):
    """libdewi_b200.so could not be loaded, or no sm_100 device is present."""


class NativeError(RuntimeError):
    """A C-ABI call returned non-zero; the message is dewi_last_error()."""


# name -> (restype, argtypes); must list every symbol of include/dewi_b200.h
SIGNATURES = {
    "dewi_abi_version": (c_int, []),
    "dewi_last_error": (c_char_p, []),
    "dewi_device_check": (c_int, [c_int, POINTER(c_int), POINTER(c_size_t), POINTER(c_size_t)]),
    "dewi_index_create": (c_int, [c_int, c_int, c_int, c_int, POINTER(c_void_p)]),
    "dewi_index_destroy": (c_int, [c_void_p]),
    "dewi_index_reserve": (c_int, [c_void_p, c_int64]),
    "dewi_index_set_id_base": (c_int, [c_void_p, c_int64]),
    "dewi_index_append": (c_int, [c_void_p, c_void_p, c_int64, c_int, c_int, c_void_p]),
    "dewi_index_set_payload": (c_int, [c_void_p, c_void_p, c_void_p, c_int64, c_int64, c_int, c_void_p]),
    "dewi_index_size": (c_int, [c_void_p, POINTER(c_int64)]),
    "dewi_index_get_row": (c_int, [c_void_p, c_int64, c_void_p]),
    "dewi_index_export_rows": (c_int, [c_void_p, c_int64, c_int64, c_void_p, c_int, c_void_p]),
    "dewi_index_export_bf16": (c_int, [c_void_p, c_int64, c_int64, c_void_p, c_int, c_void_p]),
    "dewi_index_append_bf16": (c_int, [c_void_p, c_void_p, c_int64, c_int, c_void_p]),
    "dewi_index_get_payload": (c_int, [c_void_p, c_int64, c_int64, c_void_p, c_void_p, c_int, c_void_p]),
    "dewi_index_search_local": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]),
    "dewi_index_search_local_push": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, POINTER(c_uint64), POINTER(c_uint64), c_int64, c_uint32, c_void_p]),
    "dewi_rerank_gathered": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int64, c_int, c_int, c_double, c_double, c_void_p, c_void_p, c_void_p, c_uint32, c_void_p, c_double, c_int, c_void_p]),
    "dewi_rerank": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int64, c_int, c_int, c_double, c_double, c_void_p, c_void_p, c_int, c_void_p]),
    "dewi_index_search": (c_int, [c_void_p, c_void_p, c_int, c_int, c_double, c_double, c_int, c_void_p, c_void_p, c_void_p]),
    "dewi_index_set_blend": (c_int, [c_void_p, c_double, c_double]),
    "dewi_index_last_launches": (c_int, [c_void_p, POINTER(c_int)]),
    "dewi_plan_probe": (c_int, [c_int, c_int, c_int, c_int64, c_int, c_int, c_int, c_int, c_int, POINTER(c_int)]),
    "dewi_index_cert_stats": (c_int, [c_void_p, POINTER(c_int64), POINTER(c_int64)]),
    "dewi_index_set_profiling": (c_int, [c_void_p, c_int]),
    "dewi_index_sweep_ms": (c_int, [c_void_p, c_int, POINTER(c_float), POINTER(c_int)]),
    "dewi_fit_stats": (c_int, [c_void_p, c_int64, c_int, c_int64, POINTER(c_double), POINTER(c_double), c_int, c_void_p]),
    "dewi_score": (c_int, [c_void_p, c_int, c_int64, c_int64, POINTER(c_double), POINTER(c_double), POINTER(c_double), c_int, c_void_p, c_int, c_int, c_void_p]),
    "dewi_local_weights": (c_int, [c_void_p, c_int64, c_void_p, c_int, c_void_p]),
    "dewi_cluster_pairs": (c_int, [c_void_p, c_void_p, c_int64, c_int64, c_void_p, c_int, c_void_p]),
    "dewi_similarity_dense": (c_int, [c_void_p, c_int64, c_void_p, c_int64, c_int, c_void_p, c_int, c_void_p]),
    "dewi_join_workspace_bytes": (c_int64, [c_int64, c_int64, c_int, c_int, c_int]),
    "dewi_self_join_range": (c_int, [c_void_p, c_int64, c_int, c_float, c_int64, c_int64, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int64, POINTER(c_int64), c_void_p, c_int64, c_int, c_void_p]),
    "dewi_join": (c_int, [c_void_p, c_int64, c_void_p, c_int64, c_int, c_float, c_int, c_int64, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int64, POINTER(c_int64), c_void_p, c_int64, c_int, c_void_p]),
}

_lib = None


def load_library() -> ctypes.CDLL:
    """dlopen the in-tree shared library and bind every declared symbol."""
    global _lib
    if _lib is not None:
        return _lib
    path = Path(os.environ.get("DEWI_B200_LIB", LIB_PATH))
    if not path.exists():
        raise NativeUnavailable(
            f"{path} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(nvcc, sm_100a). dewi_b200 has no CPU fallback."
        )
    try:
        lib = ctypes.CDLL(str(path))
    except OSError as e:  # pragma: no cover - depends on the machine
        raise NativeUnavailable(f"cannot load {path}: {e}") from e
    for name, (restype, argtypes) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError here means header and library disagree
        fn.restype = restype
        fn.argtypes = argtypes
    _lib = lib
    return lib


def last_error() -> str:
    msg = load_library().dewi_last_error()
    return msg.decode("utf-8", "replace") if msg else ""


def check(rc: int, exc=NativeError) -> None:
    if rc != 0:
        raise exc(last_error())


def require_device(device: int = 0) -> int:
    """Raise NativeUnavailable unless `device` is a B200; returns its SM count."""
    lib = load_library()
    sms = c_int(0)
    rc = lib.dewi_device_check(int(device), ctypes.byref(sms), None, None)
    if rc != 0:
        raise NativeUnavailable(last_error())
    return sms.value


def stream_ptr(stream=None) -> c_void_p:
    """cudaStream_t of a torch stream (default: torch's current stream)."""
    import torch

    s = stream if stream is not None else torch.cuda.current_stream()
    return c_void_p(s.cuda_stream)
