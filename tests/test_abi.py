"""The C-ABI shared library loads and exports every symbol include/viquae_b200.h declares; misuse is
reported through error codes.  No compute calls: runs without a GPU."""
import ctypes
import os
import re

import pytest

from viquae_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "viquae_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(vq_[a-z0-9_]+)\s*\(", text)))


def test_library_exists_and_loads():
    assert os.path.exists(_lib.LIB_PATH), "build with `python -m viquae_b200.build`"
    lib = _lib.load()
    assert lib.vq_version().decode().startswith("viquae_b200")


def test_every_declared_symbol_is_exported_and_bound():
    lib = ctypes.CDLL(_lib.LIB_PATH)
    names = declared_symbols()
    assert len(names) >= 20
    for name in names:
        assert hasattr(lib, name), f"{name} declared in include/viquae_b200.h but not exported"
        assert name in _lib.SIGNATURES, f"{name} has no ctypes signature in viquae_b200/_lib.py"
    assert set(_lib.SIGNATURES) == set(names)


def test_no_torch_types_in_the_header():
    text = open(os.path.join(ROOT, "include", "viquae_b200.h")).read()
    assert 'extern "C"' in text and "torch" not in text.lower().replace("pytorch", "") and "at::" not in text


def test_errors_are_codes_not_exceptions():
    lib = _lib.load()
    h = ctypes.c_void_p()
    rc = lib.vq_index_create(ctypes.byref(h), -5, _lib.BF16, 0, 0)
    assert rc == _lib.ERR_INVALID and b"bad arguments" in lib.vq_last_error()
    rc = lib.vq_index_create(ctypes.byref(h), 16, 99, 0, 0)
    assert rc == _lib.ERR_INVALID
    assert lib.vq_index_ntotal(None) == 0 and lib.vq_index_dim(None) == 0
    assert lib.vq_index_workspace_bytes(None, 10, 10) == 0
    assert lib.vq_fuse_workspace_bytes(3, 10, 100) > 0


def test_fails_loudly_without_a_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    lib = _lib.load()
    assert lib.vq_device_count() == 0
    h = ctypes.c_void_p()
    assert lib.vq_index_create(ctypes.byref(h), 16, _lib.BF16, 0, 0) < 0  # never a silent CPU path
    from viquae_b200.index import B200FlatIndex
    with pytest.raises(RuntimeError, match="no CUDA device"):
        B200FlatIndex(d=16)
