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


def test_emu_perturbed_graphs_cli(emu_lib, tmp_path):
    """nested bubbles + random chords + random-walk haplotypes through the CLI, whole VCF byte for byte vs the compiled reference"""
    import os
    n, nrec, ntang, nsubr = cc.check_perturbed_cli(os.path.join(os.path.dirname(emu_lib), "povu_emu"), str(tmp_path), range(300))
    # seed 6706: the first reference path only meets an inversion pin pair in region >1>9; the reference's trek still "has data"
    # (get_min_rov default-inserts the ref's map for every context it looks at, overlay.cpp:478-479) and goes on to the second
    # reference path, which yields a record
    assert cc.check_perturbed_cli(os.path.join(os.path.dirname(emu_lib), "povu_emu"), str(tmp_path), [6706])[1] == 3
    assert n > 250 and nrec > 3000 and ntang > 300 and nsubr > 300, (n, nrec, ntang, nsubr)
    cli = os.path.join(os.path.dirname(emu_lib), "povu_emu")
    n, nrec, ntang, nsubr = cc.check_perturbed_cli(cli, str(tmp_path), range(150), build=cc.perturbed_graph_big)
    assert n > 110 and nrec > 4000, (n, nrec, ntang, nsubr)
    # several components, two reference samples, one haplotype as the reference with chunks of three regions
    from povu_b200 import synth
    n1, r1, _, _ = cc.check_perturbed_cli(cli, str(tmp_path), range(60), build=lambda s: synth.concat_components([cc.perturbed_graph(3 * s + k) for k in range(2 + s % 2)]))
    n2, r2, _, _ = cc.check_perturbed_cli(cli, str(tmp_path), range(60), build=cc.perturbed_graph_big, args=["-P", "S001", "-P", "S003"])
    n3, r3, _, _ = cc.check_perturbed_cli(cli, str(tmp_path), range(60), args=["-P", "S002#1", "-c", "3"])
    assert n1 > 40 and n2 > 40 and n3 > 40 and r1 + r2 + r3 > 600, (n1, r1, n2, r2, n3, r3)


def test_emu_reference_class_naming(emu_lib, tmp_path):
    """Where a DFS-tree vertex has three or more children reaching above it, the reference's (top bracket, list size) class names
    merge classes that cycle equivalence keeps apart (its capping edge is too short: flubbles.cpp:564-573) and it reports flubbles
    the XOR labels do not: seeds 860 / 1058 / 2074 of the perturbed family are such graphs.  decompose detects the situation and
    re-derives the classes the reference's way (pv_refnum.cuh).  Forced on for every graph, that path must give the same results
    as the label path wherever the two agree: fixtures, LPA, synthetic families, 200 perturbed graphs."""
    import os
    import subprocess
    cli = os.path.join(os.path.dirname(emu_lib), "povu_emu")
    n, nrec, _, _ = cc.check_perturbed_cli(cli, str(tmp_path), [860, 1058, 2074])
    assert n == 3 and nrec > 40
    # the label path alone misses a flubble on those three
    for seed in (860, 1058, 2074):
        try:
            cc.check_perturbed_cli(cli, str(tmp_path), [seed], env={"POVU_B200_XOR_CLASSES": "1"})
            raise RuntimeError(f"seed {seed}: the label path alone was expected to differ from the reference")
        except AssertionError:
            pass
    force = {"POVU_B200_FAITHFUL_CLASSES": "1"}
    n, nrec, ntang, nsubr = cc.check_perturbed_cli(cli, str(tmp_path), range(200), env=force)
    assert n > 170 and nrec > 2000
    strip = lambda t: "".join(l for l in t.splitlines(True) if not l.startswith("##fileDate"))
    for name in helpers.POSITIVE + ["LPA"] + helpers.SYNTHETIC:
        prefix = {"hairpin_inversion_subr": "ref", "LPA": "chm13", "multi_comp": "S001#1#chr1"}.get(name, "S001" if name in helpers.SYNTHETIC else "HG1")
        r = subprocess.run([cli, "gfa2vcf", "-i", helpers.fixture_path(name, str(tmp_path)), "-t", "1", "-P", prefix], capture_output=True, text=True, env=dict(os.environ, **force))
        assert r.returncode == 0, (name, r.stderr[-300:])
        assert strip(r.stdout) == open(os.path.join(helpers.GOLDEN, "expected", name + ".vcf")).read(), name


def test_emu_perturbed_files_and_restrict(emu_lib, tmp_path):
    import os
    n, nrec = cc.check_perturbed_files_and_restrict(os.path.join(os.path.dirname(emu_lib), "povu_emu"), str(tmp_path), range(90))
    assert n > 70 and nrec > 1200, (n, nrec)


def test_emu_call_on_subflubble_forests(emu_lib, tmp_path):
    import os
    n, nrec, fams = cc.check_subflubble_forests(os.path.join(os.path.dirname(emu_lib), "povu_emu"), str(tmp_path), range(80))
    assert n > 60 and nrec > 3000 and fams.get("T", 0) > 3000 and fams.get("C", 0) > 20, (n, nrec, fams)
