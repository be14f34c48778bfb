"""N > 1 on real GPUs (needs >= 2 devices; skipped otherwise): components sharded over 2 NCCL ranks, merged output
byte-identical to the compiled reference."""
import os

import pytest

import helpers
from test_dist_emu import run_world

pytestmark = pytest.mark.gpu


def test_nccl_two_ranks_multi_comp(gpu, tmp_path):
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    out = str(tmp_path)
    text = run_world("cuda", helpers.fixture_path("multi_comp"), "S001#1#chr1", 100, out)
    assert text == open(os.path.join(helpers.GOLDEN, "expected", "multi_comp.vcf")).read()
    forest = os.path.join(out, "forest")
    files = {int(f[:-5]): open(os.path.join(forest, f)).read() for f in os.listdir(forest)}
    assert files == helpers.expected_pvsts("multi_comp")


def test_gpu_shard_large_graph_on_one_device(gpu, tmp_path):
    """Sharding a 24-component graph three ways on ONE device: every share is cut along component borders (the session
    verifies its device component labels against the links) and the records of every component equal those of the
    unsharded session.  Large enough for long union-find chains (a path-compression race once corrupted the labels)."""
    from povu_b200 import api, synth
    W = [248956422, 242193529, 198295559, 190214555, 181538259, 170805979, 159345973, 145138636, 138394717, 133797422, 135086622, 133275309,
         114364328, 107043718, 101991189, 90338345, 83257441, 80373285, 58617616, 64444167, 46709983, 50818468, 156040895, 57227415, 16569]
    tpl = [synth.template_from(synth.snp_indel_chain(6, n_paths=6, seed=100 + k)) for k in range(18)] + synth.cfg3_templates(seed=7, n_nested=2, tower_depth=0)
    g = synth.tiled(tpl, 3000, 8, seed=11, n_components=25, comp_weights=W)

    def records(s, prefix):
        s.call([prefix])
        return [l for l in s.vcf_text(with_date=False).splitlines() if not l.startswith("#")]

    whole = api.Session.from_graph(gpu, g)
    whole.decompose()
    seen = 0
    for rank in range(3):
        s = api.Session.from_graph(gpu, g)
        assert s.shard(rank, 3) == 24
        ga = s.graph_arrays(labels=False)
        own = [n for k, n in enumerate(ga["path_names"]) if n.startswith("S001#1#") and ga["step_off"][k + 1] > ga["step_off"][k]]
        s.hot_path([own[0]], keep_resident=True)
        for p in own[:4]:
            assert records(s, p) == records(whole, p), (rank, p)
        seen += len(own)
        s.close()
    assert seen == 24


def test_gpu_reshard_repeatable_and_record_stream(gpu):
    """The repeatable device-side partition (resident and from-host modes) and the packed binary record stream in HBM."""
    import torch
    import test_dist_emu
    test_dist_emu.check_reshard_repeatable(gpu)
    from povu_b200 import api, synth
    g = synth.concat_components([synth.snp_indel_chain(60, 6, seed=1), synth.snp_indel_chain(40, 6, seed=2)])
    s = api.Session.from_graph(gpu, g)
    s.reshard(0, 1)
    s.hot_path(["S001#1#chr1"])
    need = s.records_packed_bytes()
    buf = torch.zeros(need + 64, dtype=torch.uint8, device="cuda")
    torch.cuda.synchronize()
    assert s.records_pack(buf.data_ptr(), buf.numel()) == need
    torch.cuda.synchronize()
    h = buf.cpu().numpy()
    hdr = h[:64].view("<u8")
    nrec = int(s.lib.pvgpu_session_record_count(s.h))
    assert bytes(h[:8]) == b"PVREC001" and hdr[1] == nrec and hdr[4] == len(g["paths"])
    # left context steps carry segment ids: compare with the ids named in the VCF ID column
    import re
    R = nrec; pad = lambda x: (x + 7) & ~7
    cxtL = h[64 + 2 * pad(R * 4): 64 + 2 * pad(R * 4) + R * 4].view("<u4")
    ids_vcf = [int(re.findall(r"\d+", l.split("\t")[2])[0]) for l in s.vcf_text(with_date=False).splitlines() if not l.startswith("#")]
    assert set((cxtL >> 1).tolist()) <= set(int(x) for x in g["ids"]) and len(ids_vcf) == nrec
