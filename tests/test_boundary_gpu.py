"""CLI and povu_ffi.h boundary on the B200: the `povu` binary and libpovu_b200.so (CUDA build)."""
import os

import pytest

import boundary_checks as bc

pytestmark = pytest.mark.gpu


def test_gpu_cli_contract(gpu, tmp_path):
    bc.check_cli(os.path.join(bc.ROOT, "povu_b200", "bin", "povu"), str(tmp_path))


def test_gpu_ffi(gpu, tmp_path):
    L = bc.ffi(os.path.join(bc.ROOT, "povu_b200", "libpovu_b200.so"))
    bc.check_ffi_builder(L)
    bc.check_ffi_compute(L, str(tmp_path))


def test_gpu_export_golden(gpu, tmp_path):
    import export_checks as ec
    ec.check_export_golden(os.path.join(bc.ROOT, "povu_b200", "bin", "povu"), str(tmp_path))


def test_gpu_export_vs_reference(gpu, tmp_path):
    import export_checks as ec
    cli = os.path.join(bc.ROOT, "povu_b200", "bin", "povu")
    nb, ne = ec.check_wild_vs_reference(cli, str(tmp_path), range(14))      # (every CLI run creates a CUDA context: the 120-seed sweep is the CPU rehearsal's)
    assert nb > 4 and ne > 120
    nb, ne = ec.check_wild_vs_reference(cli, str(tmp_path), range(500, 504), V=3000)
    assert ne > 4000
    ec.check_degenerate_vs_reference(cli, str(tmp_path))
    ec.check_families_vs_reference(cli, str(tmp_path))
