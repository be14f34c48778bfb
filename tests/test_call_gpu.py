"""GPU parity tests proper for the call half: the sm_100a pipeline through the pvgpu_session_* C ABI
against the committed outputs of the compiled reference and against the CPU oracle (bit-exact text)."""
import pytest

import call_checks as cc
import helpers
from povu_b200 import api

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name", helpers.POSITIVE + helpers.SYNTHETIC + ["LPA"])
def test_gpu_call_golden(gpu, name, tmp_path):
    cc.check_golden(gpu, name, str(tmp_path))


def test_gpu_call_synthetic_vs_oracle(gpu, tmp_path):
    cc.run_all_synthetic(gpu, str(tmp_path), n_seeds=16)


def test_gpu_call_deep_regions(gpu, tmp_path):
    cc.run_deep_regions(gpu, str(tmp_path))


@pytest.mark.parametrize("name", ["nested_s1", "multi_comp", "two_ordered_substitutions"])
def test_gpu_pvst_roundtrip(gpu, name, tmp_path):
    cc.check_pvst_roundtrip(gpu, name, str(tmp_path))


@pytest.mark.parametrize("name", helpers.NEGATIVE)
def test_gpu_rejects_invalid_gfa(gpu, name):
    with pytest.raises(RuntimeError, match="Invalid GFA"):
        api.Session(gpu, helpers.fixture_path(name))


def test_gpu_call_tangled_vs_reference(gpu, tmp_path):
    cc.run_tangled_vs_reference(gpu, str(tmp_path), n_seeds=32)


def test_gpu_call_inversions_vs_reference(gpu, tmp_path):
    cc.run_inversions_vs_reference(gpu, str(tmp_path), n_seeds=60)


def test_gpu_call_cyclic_vs_reference(gpu, tmp_path):
    cc.run_cyclic_vs_reference(gpu, str(tmp_path), n_seeds=40)


@pytest.mark.parametrize("name", ["nested_s1", "hairpin_inversion_subr", "LPA"])
def test_gpu_call_hub_vertices(gpu, name, tmp_path, monkeypatch):
    """vertex -> steps index: every vertex visited more than once takes the hub path (restricted stable sort)"""
    monkeypatch.setenv("POVU_B200_VS_SMALL", "1")
    cc.check_golden(gpu, name, str(tmp_path))


@pytest.mark.parametrize("name", ["nested_s1", "LPA"])
def test_gpu_call_matfill_lanes(gpu, name, tmp_path, monkeypatch):
    """depth matrix: 32 logical lanes per column stride over the vertex's step list (the many-paths configuration)"""
    monkeypatch.setenv("POVU_B200_MATFILL_LANES", "32")
    cc.check_golden(gpu, name, str(tmp_path))


@pytest.mark.parametrize("name", ["nested_s1", "nested_s2_shuffled", "hairpin_inversion_subr", "multi_comp", "LPA"])
def test_gpu_call_many_path_kernels(gpu, name, tmp_path, monkeypatch):
    """The kernels that take over when regions have dozens of rows / candidates (hundreds of paths), forced on the fixtures:
    one warp per region for depth cells + lap 0 + context counts (C_region) and for the allele grouping (C_group)."""
    monkeypatch.setenv("POVU_B200_FUSED_REGION", "1")
    monkeypatch.setenv("POVU_B200_GROUP_WARP", "1")
    cc.check_golden(gpu, name, str(tmp_path))


def test_gpu_call_many_path_kernels_families(gpu, tmp_path, monkeypatch):
    monkeypatch.setenv("POVU_B200_FUSED_REGION", "1")
    monkeypatch.setenv("POVU_B200_GROUP_WARP", "1")
    cc.run_all_synthetic(gpu, str(tmp_path), n_seeds=8)
    cc.run_tangled_vs_reference(gpu, str(tmp_path), n_seeds=10)
    cc.run_inversions_vs_reference(gpu, str(tmp_path), n_seeds=15)


def test_gpu_perturbed_graphs(gpu, tmp_path):
    """perturbed bubble graphs with random-walk haplotypes (tests/call_checks.py), whole VCF vs the compiled reference: through
    the in-process operators (many graphs, one CUDA context) and a few through the CLI binary"""
    import os
    n, nrec = cc.check_perturbed_api(gpu, str(tmp_path), range(300, 420))
    assert n > 95 and nrec > 1000, (n, nrec)
    n, nrec = cc.check_perturbed_api(gpu, str(tmp_path), range(150, 230), build=cc.perturbed_graph_big)
    assert n > 55 and nrec > 2000, (n, nrec)
    n, nrec, _, _ = cc.check_perturbed_cli(os.path.join(cc.ROOT, "povu_b200", "bin", "povu"), str(tmp_path), range(1000, 1012))
    assert n > 8


def test_gpu_reference_class_naming(gpu, tmp_path):
    """see test_emu_reference_class_naming: graphs where the reference's class names differ from cycle equivalence, and the
    re-derived classes forced on for LPA and 80 perturbed graphs"""
    import os
    import subprocess
    cli = os.path.join(cc.ROOT, "povu_b200", "bin", "povu")
    n, nrec, _, _ = cc.check_perturbed_cli(cli, str(tmp_path), [860, 1058, 2074])
    assert n == 3 and nrec > 40
    force = {"POVU_B200_FAITHFUL_CLASSES": "1"}
    n, nrec, _, _ = cc.check_perturbed_cli(cli, str(tmp_path), range(16), env=force)
    assert n > 10 and nrec > 100
    strip = lambda t: "".join(l for l in t.splitlines(True) if not l.startswith("##fileDate"))
    for name, prefix in (("LPA", "chm13"), ("nested_s1", "S001"), ("hairpin_inversion_subr", "ref")):
        r = subprocess.run([cli, "gfa2vcf", "-i", helpers.fixture_path(name, str(tmp_path)), "-t", "1", "-P", prefix], capture_output=True, text=True, env=dict(os.environ, **force))
        assert r.returncode == 0, (name, r.stderr[-300:])
        assert strip(r.stdout) == open(os.path.join(helpers.GOLDEN, "expected", name + ".vcf")).read(), name
