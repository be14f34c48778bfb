"""The C-ABI shared library loads on a machine without a GPU and exports every symbol include/*.h declares."""
import ctypes
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared(header):
    txt = open(os.path.join(ROOT, "include", header)).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(pvgpu_[a-z0-9_]+|povu_[a-z0-9_]+)\s*\(", txt)))


def test_povu_b200_h_symbols_exported(built):
    built.build_cuda()
    lib = ctypes.CDLL(os.path.join(ROOT, "povu_b200", "libpovu_b200.so"))
    names = _declared("povu_b200.h")
    assert len(names) >= 15
    for n in names:
        assert hasattr(lib, n), f"missing export {n}"
    assert lib.pvgpu_build_kind() == 1


def test_povu_ffi_h_symbols_exported(built):
    built.build_cuda()
    lib = ctypes.CDLL(os.path.join(ROOT, "povu_b200", "libpovu_b200.so"))
    names = _declared("povu_ffi.h")
    assert len(names) == 33          # povu_ffi.h:147-437 declares 33 functions
    for n in names:
        assert hasattr(lib, n), f"missing export {n}"


def test_create_fails_loudly_without_gpu(built):
    """No CPU fallback: without a CUDA device pvgpu_create must fail (skipped on a GPU box)."""
    import pytest
    try:
        import torch
        if torch.cuda.is_available():
            pytest.skip("GPU present")
    except ImportError:
        pass
    from povu_b200 import api
    with pytest.raises(RuntimeError):
        api.Povu(device=0)
