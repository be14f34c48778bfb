"""Call half on the host debug rehearsal (tests/emu/libpovu_emu.so: the SAME kernel bodies, one logical
thread at a time, three thread orders) -- validates the pipeline logic where no GPU exists.  The GPU
parity tests proper are in test_call_gpu.py."""
import os

import pytest

import call_checks as cc
import helpers
from povu_b200 import api


def _ctx(emu_lib, order):
    os.environ["PV_EMU_ORDER"] = str(order)
    return api.Povu(lib_path=emu_lib)


@pytest.mark.parametrize("name", helpers.POSITIVE + helpers.SYNTHETIC + ["LPA"])
def test_emu_call_golden(emu_lib, name, tmp_path):
    cc.check_golden(_ctx(emu_lib, 0), name, str(tmp_path))


@pytest.mark.parametrize("order", [0, 1, 2])
def test_emu_call_synthetic_vs_oracle(emu_lib, built, order, tmp_path):
    cc.run_all_synthetic(_ctx(emu_lib, order), str(tmp_path), n_seeds=8)


def test_emu_call_deep_regions(emu_lib, built, tmp_path):
    cc.run_deep_regions(_ctx(emu_lib, 2), str(tmp_path))


@pytest.mark.parametrize("name", ["nested_s1", "multi_comp", "two_ordered_substitutions"])
def test_emu_pvst_roundtrip(emu_lib, name, tmp_path):
    cc.check_pvst_roundtrip(_ctx(emu_lib, 0), name, str(tmp_path))


@pytest.mark.parametrize("name", helpers.NEGATIVE)
def test_emu_rejects_invalid_gfa(emu_lib, name):
    pv = _ctx(emu_lib, 0)
    with pytest.raises(RuntimeError, match="Invalid GFA"):
        api.Session(pv, helpers.fixture_path(name))


def test_emu_call_tangled_vs_reference(emu_lib, built, tmp_path):
    cc.run_tangled_vs_reference(_ctx(emu_lib, 1), str(tmp_path), n_seeds=16)


def test_emu_call_inversions_vs_reference(emu_lib, built, tmp_path):
    cc.run_inversions_vs_reference(_ctx(emu_lib, 2), str(tmp_path), n_seeds=30)


def test_emu_call_cyclic_vs_reference(emu_lib, built, tmp_path):
    cc.run_cyclic_vs_reference(_ctx(emu_lib, 1), str(tmp_path), n_seeds=24)


@pytest.mark.parametrize("name", ["nested_s1", "hairpin_inversion_subr", "LPA"])
def test_emu_call_hub_vertices(emu_lib, name, tmp_path, monkeypatch):
    """vertex -> steps index: every vertex visited more than once takes the hub path (restricted stable sort)"""
    monkeypatch.setenv("POVU_B200_VS_SMALL", "1")
    cc.check_golden(_ctx(emu_lib, 2), name, str(tmp_path))


@pytest.mark.parametrize("name", ["nested_s1", "LPA"])
def test_emu_call_matfill_lanes(emu_lib, name, tmp_path, monkeypatch):
    """depth matrix: 32 logical lanes per column stride over the vertex's step list (the many-paths configuration)"""
    monkeypatch.setenv("POVU_B200_MATFILL_LANES", "32")
    cc.check_golden(_ctx(emu_lib, 2), name, str(tmp_path))
