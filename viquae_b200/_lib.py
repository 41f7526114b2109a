"""ctypes binding of libviquae_b200.so — the C ABI declared in include/viquae_b200.h.

There is no fallback of any kind: if the shared library is missing or a call fails, this raises.
"""
from __future__ import annotations

import ctypes
import os
from ctypes import POINTER, c_char_p, c_double, c_float, c_int, c_int32, c_int64, c_size_t, c_uint64, c_void_p

_HERE = os.path.dirname(os.path.abspath(__file__))
# VIQUAE_B200_LIB points at another build of the same ABI (A/B runs of a kernel change on one box)
LIB_PATH = os.environ.get("VIQUAE_B200_LIB") or os.path.join(_HERE, "libviquae_b200.so")

F32, BF16, F16 = 0, 1, 2
F32X3 = 3  # storage only: fp32 values split into three bf16 terms, tensor-core contraction at fp32 fidelity
VARIANT_AUTO, VARIANT_TC_CG1, VARIANT_TC_CG2 = 0, 1, 2
ERR_CUDA, ERR_INVALID, ERR_UNSUPPORTED, ERR_NOMEM = -1, -2, -3, -4

# every symbol include/viquae_b200.h declares: name -> (restype, argtypes)
SIGNATURES = {
    "vq_device_count": (c_int, []),
    "vq_last_error": (c_char_p, []),
    "vq_version": (c_char_p, []),
    "vq_index_create": (c_int, [POINTER(c_void_p), c_int, c_int, c_int, c_uint64]),
    "vq_index_destroy": (None, [c_void_p]),
    "vq_index_add": (c_int, [c_void_p, c_void_p, c_uint64, c_int, c_int, c_void_p]),
    "vq_index_ntotal": (c_uint64, [c_void_p]),
    "vq_index_dim": (c_int, [c_void_p]),
    "vq_index_dtype": (c_int, [c_void_p]),
    "vq_index_set_id_offset": (c_int, [c_void_p, c_int64]),
    "vq_index_set_variant": (c_int, [c_void_p, c_int]),
    "vq_index_set_timing": (c_int, [c_void_p, c_int]),
    "vq_index_read_timing": (c_int, [c_void_p, POINTER(c_double), POINTER(c_int)]),
    "vq_index_launch_count": (c_uint64, [c_void_p]),
    "vq_index_copy_rows": (c_int, [c_void_p, c_uint64, c_uint64, c_void_p, c_int, c_void_p]),
    "vq_index_workspace_bytes": (c_size_t, [c_void_p, c_int, c_int]),
    "vq_index_search": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p, c_size_t, c_void_p]),
    "vq_index_search_host": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_void_p, c_void_p]),
    "vq_prep_rows": (c_int, [c_void_p, c_int, c_void_p, c_int, c_uint64, c_int, c_int, c_void_p]),
    "vq_merge_topk": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p]),
    "vq_merge_topk_strided": (c_int, [c_void_p, c_void_p, c_int64, c_int64, c_int, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p]),
    "vq_xchg_create": (c_int, [POINTER(c_void_p), c_int, c_int, c_int, c_size_t]),
    "vq_xchg_export": (c_int, [c_void_p, c_void_p]),
    "vq_xchg_attach": (c_int, [c_void_p, c_int, c_void_p]),
    "vq_xchg_slot_bytes": (c_size_t, [c_void_p]),
    "vq_xchg_publish": (c_int, [c_void_p, c_void_p, c_size_t, c_void_p]),
    "vq_xchg_merge_topk": (c_int, [c_void_p, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p]),
    "vq_xchg_gather": (c_int, [c_void_p, c_size_t, c_void_p, c_void_p]),
    "vq_xchg_destroy": (None, [c_void_p]),
    "vq_expand_hits": (c_int, [c_void_p, c_void_p, c_int, c_int, c_void_p, c_void_p, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p,
                       c_void_p, c_void_p]),
    "vq_fuse_workspace_bytes": (c_size_t, [c_int, c_int, c_int]),
    "vq_fuse": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p,
                        c_void_p, c_size_t, c_void_p]),
    "vq_inbatch_logprobs": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_int64, c_void_p, c_void_p, c_void_p, c_void_p]),
    "vq_mark_relevant": (c_int, [c_void_p, c_int, c_int, c_void_p, c_void_p, c_void_p, c_void_p]),
    "vq_fuse_grid_workspace_bytes": (c_size_t, [c_int, c_int, c_int, c_int]),
    "vq_fuse_grid": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_void_p, c_int, c_void_p, c_int, c_void_p, c_int, c_int, c_int,
                             c_void_p, c_void_p, c_void_p, c_size_t, c_void_p]),
}

_lib = None


class EngineError(RuntimeError):
    """A C-ABI call returned a negative code (the reference surfaces FAISS C++ exceptions as RuntimeError)."""

    def __init__(self, code: int, message: str):
        super().__init__(f"viquae_b200 error {code}: {message}")
        self.code = code


def load() -> ctypes.CDLL:
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(
                f"{LIB_PATH} is missing: build it with `python -m viquae_b200.build` (nvcc, sm_100a). "
                "viquae_b200 has no CPU or PyTorch fallback."
            )
        lib = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(lib, name)  # AttributeError if the library does not export a declared symbol
            fn.restype = res
            fn.argtypes = args
        _lib = lib
    return _lib


def check(rc: int) -> None:
    if rc != 0:
        raise EngineError(rc, load().vq_last_error().decode("utf-8", "replace"))
