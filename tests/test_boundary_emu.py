"""CLI and povu_ffi.h boundary on the host rehearsal build (no GPU here); the GPU versions are in
test_boundary_gpu.py.  Also: the CUDA library exports every povu_ffi.h symbol."""
import os

import boundary_checks as bc


def test_emu_cli_contract(emu_lib, built, tmp_path):
    bc.check_cli(os.path.join(os.path.dirname(emu_lib), "povu_emu"), str(tmp_path))


def test_emu_ffi_builder(emu_lib):
    bc.check_ffi_builder(bc.ffi(emu_lib))


def test_emu_ffi_compute(emu_lib, tmp_path):
    bc.check_ffi_compute(bc.ffi(emu_lib), str(tmp_path))


def test_cuda_library_ffi_builder_without_gpu(built):
    """Graph building / accessors need no device, so they work on the CUDA build here too."""
    built.build_cuda()
    bc.check_ffi_builder(bc.ffi(os.path.join(bc.ROOT, "povu_b200", "libpovu_b200.so")))
