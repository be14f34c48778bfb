"""Device VCF text writer on the host debug rehearsal (same kernel bodies, compiled for the CPU): see text_checks.py."""
import os

import text_checks as tc
from povu_b200 import api


def _ctx(emu_lib, order=0):
    os.environ["PV_EMU_ORDER"] = str(order)
    return api.Povu(lib_path=emu_lib)


def test_emu_af_and_integers(emu_lib):
    tc.check_af_and_integers(_ctx(emu_lib))


def test_emu_shared_slots_and_prefix_files(emu_lib, built, tmp_path):
    tc.check_shared_slots_and_prefix_files(_ctx(emu_lib, 2), str(tmp_path))


def test_emu_fd_write(emu_lib, built, tmp_path):
    tc.check_fd_write(_ctx(emu_lib), str(tmp_path))
