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
    assert bc.check_ffi_perturbed(bc.ffi(emu_lib), str(tmp_path), range(60)) > 45


def test_cuda_library_ffi_builder_without_gpu(built):
    """Graph building / accessors need no device, so they work on the CUDA build here too."""
    built.build_cuda()
    bc.check_ffi_builder(bc.ffi(os.path.join(bc.ROOT, "povu_b200", "libpovu_b200.so")))


def test_emu_export_golden(emu_lib, tmp_path):
    """--structure-export / --hairpins against the committed outputs of the compiled reference"""
    import export_checks as ec
    ec.check_export_golden(os.path.join(os.path.dirname(emu_lib), "povu_emu"), str(tmp_path))
    ec.check_cycle_tables(os.path.join(os.path.dirname(emu_lib), "povu_emu"), str(tmp_path))
    ec.check_hierarchy_spans(os.path.join(os.path.dirname(emu_lib), "povu_emu"), str(tmp_path))
    ec.check_call_export(os.path.join(os.path.dirname(emu_lib), "povu_emu"), str(tmp_path))


def test_emu_export_vs_reference(emu_lib, tmp_path):
    """the same two surfaces against the compiled reference itself: 120 arbitrary bidirected graphs (hairpin boundaries, class
    ids, edge ids) and graph families with real variant calls, incl. LPA"""
    import export_checks as ec
    cli = os.path.join(os.path.dirname(emu_lib), "povu_emu")
    nb, ne = ec.check_wild_vs_reference(cli, str(tmp_path), range(120))
    assert nb > 60 and ne > 1500
    nb, ne = ec.check_wild_vs_reference(cli, str(tmp_path), range(500, 506), V=600)
    assert ne > 1500
    ec.check_degenerate_vs_reference(cli, str(tmp_path))
    ec.check_families_vs_reference(cli, str(tmp_path))
