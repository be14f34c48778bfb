"""N > 1 path on CPU: two gloo ranks, components sharded by the LPT rule, record lines gathered to rank 0.
The merged output must be byte-identical to the single-process output and to the compiled reference."""
import os
import socket
import subprocess
import sys

import pytest

import call_checks as cc
import helpers
from povu_b200 import api, synth

HERE = os.path.dirname(os.path.abspath(__file__))


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close()
    return p


def run_world(lib, gfa, prefix, chunk, out, world=2):
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr", "127.0.0.1",
           "--master-port", str(_free_port()), os.path.join(HERE, "dist_worker.py"), lib, gfa, prefix, str(chunk), out]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    return open(os.path.join(out, "out.vcf")).read()


def test_gloo_two_ranks_multi_comp_matches_reference(emu_lib, tmp_path):
    out = str(tmp_path)
    text = run_world(emu_lib, helpers.fixture_path("multi_comp"), "S001#1#chr1", 100, out)
    assert text == open(os.path.join(helpers.GOLDEN, "expected", "multi_comp.vcf")).read()
    forest = os.path.join(out, "forest")
    files = {int(f[:-5]): open(os.path.join(forest, f)).read() for f in os.listdir(forest)}
    assert files == helpers.expected_pvsts("multi_comp")


@pytest.mark.parametrize("chunk", [100, 3])
def test_gloo_two_ranks_equals_single_process(emu_lib, built, tmp_path, chunk):
    g = synth.concat_components([synth.nested_bubbles(71, n_top=6), synth.snp_indel_chain(25, 4, seed=8), synth.nested_bubbles(73, n_top=3),
                                 synth.snp_indel_chain(40, 4, seed=9), synth.nested_bubbles(75, n_top=8, max_depth=4)])
    gfa = os.path.join(str(tmp_path), "g.gfa")
    synth.write_gfa(g, gfa)
    pv = api.Povu(lib_path=emu_lib)
    single, _, _ = cc.gfa2vcf(pv, gfa, ["S001"], chunk)
    text = run_world(emu_lib, gfa, "S001", chunk, str(tmp_path))
    assert text == single
    assert cc.split(text, False)[1] == cc.split(cc.oracle_vcf(gfa, ["S001"], chunk), True)[1]


def test_shard_balance_and_coverage(emu_lib, tmp_path):
    """Every component is owned by exactly one rank; shares are balanced by the LPT rule."""
    g = synth.concat_components([synth.snp_indel_chain(n, 2, seed=n) for n in (50, 10, 30, 20, 40, 5)])
    gfa = os.path.join(str(tmp_path), "g.gfa")
    synth.write_gfa(g, gfa)
    pv = api.Povu(lib_path=emu_lib)
    whole = api.Session(pv, gfa)
    V, E, H, S = whole.sizes()
    tot = [0, 0, 0]; loads = []
    for r in range(3):
        s = api.Session(pv, gfa)
        assert s.shard(r, 3) == 6
        v, e, h, st = s.sizes()
        assert h == H
        tot[0] += v; tot[1] += e; tot[2] += st
        loads.append(v + e + st)
    assert tot == [V, E, S]
    assert max(loads) <= 1.35 * (sum(loads) / 3)


def check_reshard_repeatable(pv, n_ranks=3):
    """pvgpu_session_reshard on ONE session, again and again: every share equals the share of a one-shot shard(), whether the
    whole graph is still in HBM or comes from the page-locked host arrays again (then only the owned paths' walks travel);
    the binary record stream of a share has the advertised size."""
    g = synth.concat_components([synth.nested_bubbles(81, n_top=5), synth.snp_indel_chain(20, 4, seed=3), synth.nested_bubbles(83, n_top=4, max_depth=2),
                                 synth.snp_indel_chain(35, 4, seed=4), synth.snp_indel_chain(12, 4, seed=5)])
    s = api.Session.from_graph(pv, g)

    def share_vcfs(sess):
        ga = sess.graph_arrays(labels=False)
        own = [n for k, n in enumerate(ga["path_names"]) if n.startswith("S001#1#") and ga["step_off"][k + 1] > ga["step_off"][k]]
        sess.decompose()
        out = {}
        for p in own:
            sess.call([p])
            out[p] = sess.vcf_text(with_date=False)
        return out, ga
    want = []
    for r in range(n_ranks):
        one = api.Session.from_graph(pv, g)
        assert one.shard(r, n_ranks) == 5
        want.append(share_vcfs(one))
    for from_host in (False, True, True, False):
        for r in range(n_ranks):
            n, _ = s.reshard(r, n_ranks, from_host=from_host)
            assert n == 5
            got, ga = share_vcfs(s)
            assert got == want[r][0], (from_host, r)
            for k in ("ids", "ep1", "ep2", "steps", "step_off"):
                assert (ga[k] == want[r][1][k]).all(), (k, from_host, r)
    assert s.records_packed_bytes() >= 64
    assert sum(len(w[0]) for w in want) == 5


def test_emu_reshard_repeatable(emu_lib):
    check_reshard_repeatable(api.Povu(lib_path=emu_lib))
