"""ctypes binding of the C ABI declared in ``include/rvq_b200.h``.

The product path has no fallback: if ``librvq_b200.so`` is missing or a call
fails, a ``RuntimeError`` is raised.  Nothing here imports ``oracle/``.
"""
from __future__ import annotations

import ctypes as C
import os
import typing as tp

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "librvq_b200.so")

FLAG_STE = 1
FLAG_FORCE_EXACT = 2
FLAG_DIRECT_DIST = 4
FLAG_ACCUM_Q = 8
FLAG_CODES_BKT = 16
FLAG_OUT_BDT = 32

_lib: tp.Optional[C.CDLL] = None

_vp, _i, _i64, _sz, _dbl, _f = C.c_void_p, C.c_int, C.c_int64, C.c_size_t, C.c_double, C.c_float

# name -> (restype, argtypes); must list every symbol include/rvq_b200.h declares
SIGNATURES: tp.Dict[str, tp.Tuple[tp.Any, tp.List[tp.Any]]] = {
    "rvq_version": (_i, []),
    "rvq_last_error": (C.c_char_p, []),
    "rvq_device_ok": (_i, []),
    "rvq_launch_count": (C.c_uint64, []),
    "rvq_pack_bytes": (_sz, [_i, _i, _i]),
    "rvq_pack": (_i, [_vp, _i, _i, _i, _vp, _sz, _vp]),
    "rvq_encode": (_i, [_vp, _i, _i, _vp, _i64, _i64, _i64, _i, _i, _i, _i, _vp, _vp, _vp, _vp, _i, _vp]),
    "rvq_encode_train": (_i, [_vp, _i, _i, _vp, _i64, _i64, _i64, _i, _i, _i, _i, _vp, _vp, _vp, _vp, _vp, _vp, _i, _vp]),
    "rvq_decode": (_i, [_vp, _i, _i, _vp, _i64, _i64, _i64, _i, _i, _i, _vp, _vp]),
    "rvq_decode_ex": (_i, [_vp, _i, _i, _vp, _i64, _i64, _i64, _i, _i, _i, _vp, _i, _vp]),
    "rvq_ema_stats": (_i, [_vp, _i, _i, _vp, _i64, _i64, _i64, _i, _i, _i, _i, _vp, _vp, _vp, _i, _vp]),
    "rvq_ema_apply": (_i, [_vp, _vp, _vp, _i, _i, _i, _vp, _vp, _dbl, _dbl, _vp]),
    "rvq_expire_replace": (_i, [_vp, _vp, _vp, _i, _i, _f, _vp]),
    "rvq_expire_codes": (_i, [_vp, _i, _i, _vp, _i64, _i64, _i64, _i, _i, _i, _i, _vp, _vp, _vp, _f, _vp, _i, _vp]),
    "rvq_expire_stack": (_i, [_vp, _i, _i, _vp, _i64, _i64, _i64, _i, _i, _i, _i, _vp, _vp, _vp, _f, C.c_uint64, C.c_uint64,
                               _vp, _vp, _i, _vp]),
    "rvq_bitpack_bytes": (_sz, [_i, _i, _i]),
    "rvq_bitpack": (_i, [_vp, _i64, _i64, _i64, _i, _i, _i, _i, _vp, _i64, _vp]),
    "rvq_bitunpack": (_i, [_vp, _i64, _i, _i, _i, _i, _vp, _i64, _i64, _i64, _vp]),
    "rvq_kmeans_assign": (_i, [_vp, _i, _i, _vp, _i64, _vp, _vp]),
    "rvq_kmeans_update": (_i, [_vp, _i64, _i, _vp, _i, _vp, _vp, _vp, _vp]),
    "rvq_residual_combine": (_i, [_vp, _i, _i, _vp, _i64, _i64, _i64, _i, _i, _i, _i, _vp, _vp, _vp, _i, _vp]),
    "rvq_search_counters": (_i, [_vp]),
    "rvq_pack_bound_mode": (_i, [_i]),
}


def load() -> C.CDLL:
    """Load the library (once).  Raises if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f"{LIB_PATH} not found: build it with `python -m encodec_pytorch_b200.build` "
            "(there is no CPU or PyTorch fallback for the RVQ path)")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    if lib.rvq_version() != 1:
        raise RuntimeError("librvq_b200.so ABI version mismatch")
    _lib = lib
    return lib


def last_error() -> str:
    return load().rvq_last_error().decode("utf-8", "replace")


def check(rc: int, what: str) -> None:
    if rc != 0:
        raise RuntimeError(f"{what} failed (code {rc}): {last_error()}")


def stream_ptr(device: torch.device) -> int:
    return torch.cuda.current_stream(device).cuda_stream


def ptr(t: tp.Optional[torch.Tensor]) -> tp.Optional[int]:
    return None if t is None else t.data_ptr()


def ptr_array(tensors: tp.Sequence[torch.Tensor]):
    arr = (C.c_void_p * max(1, len(tensors)))()
    for i, t in enumerate(tensors):
        arr[i] = t.data_ptr()
    return arr


def require_cuda_f32(x: torch.Tensor, what: str) -> None:
    if not x.is_cuda:
        raise RuntimeError(f"{what}: expected a CUDA tensor; the B200 RVQ path has no CPU fallback")
    if x.dtype != torch.float32:
        raise RuntimeError(f"{what}: expected float32 (the reference mm raises on {x.dtype} too)")


def launch_count() -> int:
    return int(load().rvq_launch_count())
