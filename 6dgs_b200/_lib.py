"""ctypes binding of libsixdgs.so (the C ABI declared in include/sixdgs.h).

There is no CPU fallback: every op in this package goes through the CUDA library.  If the shared
object is missing, or a tensor is not a contiguous CUDA tensor of the expected dtype, the call
fails loudly.
"""
from __future__ import annotations

import ctypes
import os
from typing import Optional

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "csrc", "libsixdgs.so")
F32, BF16, F16X2, F16F8 = 0, 1, 2, 3
FEAT = 384
MAX_TOKENS = 256

_lib = None

c_p = ctypes.c_void_p
c_i = ctypes.c_int
c_i64 = ctypes.c_int64
c_sz = ctypes.c_size_t

_SIGNATURES = {
    "sixdgs_version": ([], c_i),
    "sixdgs_last_error": ([], ctypes.c_char_p),
    "sixdgs_device_supported": ([], c_i),
    "sixdgs_degrade_mask": ([c_p, c_i64, c_i, c_p, c_p, c_p], c_i),
    "sixdgs_knn_normals": ([c_p, c_i64, c_i64, c_i64, c_i, c_p, c_p], c_i),
    "sixdgs_knn_grid_workspace": ([c_i64, c_i64], c_sz),
    "sixdgs_knn_normals_grid": ([c_p, c_i64, c_i64, c_i64, c_i, c_p, ctypes.c_float, c_p, c_p, c_p, c_sz, c_p], c_i),
    "sixdgs_sym_eig3x3": ([c_p, c_i64, ctypes.c_float, c_p, c_p, c_p], c_i),
    "sixdgs_raygen_cells": ([c_p, c_p, c_i64, c_i, c_p, c_p], c_i),
    "sixdgs_raygen_fill": ([c_p, c_p, c_p, c_p, c_i, c_i, c_p, c_i64, c_p, c_i, c_i, c_i, c_p, c_p, c_p, c_p, c_p, c_p, c_p], c_i),
    "sixdgs_raygen_compact": ([c_p, c_p, c_p, c_i64, c_p, c_p, c_p, c_p, c_p, c_p, c_p, c_p, c_p], c_i),
    "sixdgs_exclusive_scan": ([c_p, c_i64, c_p, c_p], c_i),
    "sixdgs_ray_features_workspace": ([c_i64], c_sz),
    "sixdgs_ray_features": ([c_p, c_p, c_p, c_i64] + [c_p] * 10 + [c_p, c_i, c_p, c_i, c_p, c_sz, c_p], c_i),
    "sixdgs_ray_features_x2_workspace": ([c_i64], c_sz),
    "sixdgs_ray_features_x2": ([c_p, c_p, c_p, c_i64] + [c_p] * 10 + [c_p, c_p, c_p, c_sz, c_p], c_i),
    "sixdgs_linear": ([c_p, c_i64, c_i, c_i, c_p, c_p, c_i, c_p, c_i, c_i, c_p], c_i),
    "sixdgs_score_parts": ([c_i], c_i),
    "sixdgs_score_workspace": ([c_i], c_sz),
    "sixdgs_score_pass1": ([c_p, c_i, c_i64, c_p, c_i, c_p, c_p, c_i, c_p, c_sz, c_p], c_i),
    "sixdgs_score_merge": ([c_p, c_p, c_i, c_i, c_i64, c_i, c_p, c_p, c_p, c_p], c_i),
    "sixdgs_score_pass2": ([c_p, c_i, c_i64, c_p, c_i, c_p, c_p, c_p, c_p, c_i, c_p, c_sz, c_p], c_i),
    "sixdgs_score_batch_max": ([], c_i),
    "sixdgs_score_batch_parts": ([], c_i),
    "sixdgs_score_batch_workspace": ([c_i], c_sz),
    "sixdgs_score_pass1_batch": ([c_p, c_i, c_i64, c_p, c_i, c_i, c_p, c_p, c_p, c_sz, c_p], c_i),
    "sixdgs_score_pass2_batch": ([c_p, c_i, c_i64, c_p, c_i, c_i, c_p, c_p, c_p, c_i64, c_p, c_sz, c_p], c_i),
    "sixdgs_score_backward_parts": ([], c_i),
    "sixdgs_score_backward_gbar": ([c_p, c_i64, c_p, c_i, c_p, c_p, c_p, c_p, c_p, c_p], c_i),
    "sixdgs_score_backward_dlogits": ([c_p, c_i64, c_p, c_i, c_p, c_p, c_p, c_p, c_p, c_p, c_i64, c_p], c_i),
    "sixdgs_split_keys": ([c_p, c_i64, c_p, c_p, c_p], c_i),
    "sixdgs_keys_f16x2_to_f16f8": ([c_p, c_i64, c_p], c_i),
    "sixdgs_ls_partial_rows": ([], c_i),
    "sixdgs_score_pass2_batch_ls": ([c_p, c_i, c_i64, c_p, c_i, c_i, c_p, c_p, c_p, c_i64, c_p, c_p, c_p, c_p, c_p, c_sz, c_p], c_i),
    "sixdgs_ls_solve": ([c_p, c_i, ctypes.c_double, c_p, c_p, c_p, c_p, c_p, c_p, c_p], c_i),
    "sixdgs_topk_workspace": ([c_i64, c_i], c_sz),
    "sixdgs_topk": ([c_p, c_i64, c_i, c_p, c_p, c_p, c_sz, c_p], c_i),
    "sixdgs_line_intersect": ([c_p, c_p, c_p, c_i64, c_p, c_p, c_p, c_p], c_i),
    "sixdgs_pose_tail": ([c_p, c_p, c_i64, c_p, c_p, c_i, c_p, c_p, c_p, c_p], c_i),
    "sixdgs_gather_candidates": ([c_p, c_p, c_i, c_i, c_p, c_p, c_p, c_p], c_i),
}

EXPORTED_SYMBOLS = tuple(_SIGNATURES)


class SixdgsError(RuntimeError):
    pass


def load(path: Optional[str] = None) -> ctypes.CDLL:
    """dlopen the library and attach argument types.  No device is touched."""
    global _lib
    if _lib is not None and path is None:
        return _lib
    p = path or LIB_PATH
    if not os.path.exists(p):
        raise SixdgsError(
            f"{p} not found: the CUDA extension is required (no CPU fallback). "
            "Build it with `python 6dgs_b200/csrc/build.py` or `__graft_entry__.build()`.")
    lib = ctypes.CDLL(p)
    for name, (args, res) in _SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError here means the ABI and the header diverged
        fn.argtypes = args
        fn.restype = res
    if path is None:
        _lib = lib
    return lib


def call(name: str, *args):
    lib = load()
    rc = getattr(lib, name)(*args)
    if rc != 0:
        raise SixdgsError(f"{name} failed ({rc}): {lib.sixdgs_last_error().decode()}")


def stream_ptr() -> int:
    return torch.cuda.current_stream().cuda_stream


def dptr(t: Optional[torch.Tensor], dtype: Optional[torch.dtype] = torch.float32, name: str = "tensor") -> Optional[int]:
    """device pointer of a contiguous CUDA tensor (None passes through as NULL)."""
    if t is None:
        return None
    if not t.is_cuda:
        raise SixdgsError(f"{name} must be a CUDA tensor (this package has no CPU path); got device {t.device}")
    if dtype is not None and t.dtype != dtype:
        raise SixdgsError(f"{name} must be {dtype}, got {t.dtype}")
    if not t.is_contiguous():
        raise SixdgsError(f"{name} must be contiguous")
    return t.data_ptr()


def f32c(t: torch.Tensor) -> torch.Tensor:
    """contiguous fp32 view/copy on the same (CUDA) device."""
    return t.detach().to(torch.float32).contiguous()
