"""Device GFA parser on the B200 (pv_gfa.cuh): identical arrays / error messages to the host loader and the Python reader."""
import pytest

import gfa_checks as gc

pytestmark = pytest.mark.gpu


def test_gpu_gfa_fixtures(gpu, tmp_path):
    gc.check_fixtures(gpu, str(tmp_path))


def test_gpu_gfa_edge_cases_and_errors(gpu, tmp_path):
    gc.check_edge_cases(gpu, str(tmp_path))
    gc.check_errors(gpu, str(tmp_path))


def test_gpu_gfa_fuzz(gpu, tmp_path):
    gc.check_fuzz(gpu, str(tmp_path), n_cases=300)


def test_gpu_gfa_chain_200k(gpu, tmp_path):
    gc.check_large_synthetic(gpu, str(tmp_path), n_bubbles=200_000)
