"""Device GFA parser on the host debug rehearsal (same kernel bodies compiled for the CPU, three thread orders): see gfa_checks.py."""
import os

import pytest

import gfa_checks as gc
from povu_b200 import api


def _ctx(emu_lib, order=0):
    os.environ["PV_EMU_ORDER"] = str(order)
    return api.Povu(lib_path=emu_lib)


@pytest.mark.parametrize("order", [0, 1, 2])
def test_emu_gfa_fixtures(emu_lib, built, tmp_path, order):
    gc.check_fixtures(_ctx(emu_lib, order), str(tmp_path))


@pytest.mark.parametrize("order", [0, 2])
def test_emu_gfa_edge_cases_and_errors(emu_lib, built, tmp_path, order):
    pv = _ctx(emu_lib, order)
    gc.check_edge_cases(pv, str(tmp_path))
    gc.check_errors(pv, str(tmp_path))


def test_emu_gfa_fuzz(emu_lib, built, tmp_path):
    gc.check_fuzz(_ctx(emu_lib, 1), str(tmp_path), n_cases=120)


def test_emu_gfa_chain(emu_lib, built, tmp_path):
    gc.check_large_synthetic(_ctx(emu_lib, 2), str(tmp_path), n_bubbles=5000)
