import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def built():
    import __graft_entry__ as ge
    ge.build_oracle()
    return ge


@pytest.fixture(scope="session")
def emu_lib(built):
    return built.build_emu()


@pytest.fixture(scope="session")
def gpu(built):
    """A povu_b200 context on cuda:0 -- the CUDA library must load and a device must exist."""
    from povu_b200 import api
    built.build_cuda()
    return api.Povu(device=0)
