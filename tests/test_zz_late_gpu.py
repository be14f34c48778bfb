"""GPU checks added after the round's last B200 session (the box budget was spent): the same code paths as the tests that ran
there, on more inputs.  The file sorts last so that `pytest -x` reaches every test that was seen green on the B200 first."""
import os

import pytest

import boundary_checks as bc
import call_checks as cc
import export_checks as ec

pytestmark = pytest.mark.gpu
CLI = os.path.join(bc.ROOT, "povu_b200", "bin", "povu")


def test_gpu_hub_graphs_reference_classes(gpu, tmp_path):
    """hubs + chords on nested bubbles: most of these take the classes re-derived the reference's way (call_checks.hub_graph)"""
    n, faithful, nref = cc.check_hub_graphs(gpu, str(tmp_path), range(240, 400))
    assert n == 160 and faithful > 80 and nref > 30, (n, faithful, nref)


def test_gpu_cycle_tables_and_hierarchy_spans(gpu, tmp_path):
    """the conformance harness's expectations (main.rs:1490-1622) on the CUDA build's structure export"""
    ec.check_cycle_tables(CLI, str(tmp_path))
    ec.check_hierarchy_spans(CLI, str(tmp_path))
    ec.check_call_export(CLI, str(tmp_path))


def test_gpu_ffi_builder_on_perturbed_graphs(gpu, tmp_path):
    L = bc.ffi(os.path.join(bc.ROOT, "povu_b200", "libpovu_b200.so"))
    assert bc.check_ffi_perturbed(L, str(tmp_path), range(60, 100)) > 30


def test_gpu_has_data_after_pin_pair_only(gpu, tmp_path):
    """perturbed seed 6706 (see test_emu_perturbed_graphs_cli): a region whose first reference path only yields an inversion pin pair
    still goes on to the second reference path"""
    assert cc.check_perturbed_cli(CLI, str(tmp_path), [6706])[1] == 3


def test_gpu_call_on_subflubble_forests(gpu, tmp_path):
    """see test_emu_call_on_subflubble_forests (host-side forest handling; a few graphs through the CUDA CLI)"""
    n, nrec, fams = cc.check_subflubble_forests(CLI, str(tmp_path), range(12), with_lpa=True)
    assert n > 8 and nrec > 2000, (n, nrec, fams)
