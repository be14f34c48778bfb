"""Shared parity checks for the call half (used by test_call_emu.py on CPU and test_call_gpu.py on the
B200): VCF text produced through the pvgpu_session_* C ABI against (a) the committed outputs of the
compiled reference (tests/golden/expected) and (b) the CPU oracle (oracle/povu_oracle) on seeded synthetic
graphs, and (c) the compiled reference itself (oracle/_ref) for tangled regions and the SNE pass.  Bit-exact (text)."""
import os
import subprocess
import tempfile

import numpy as np

import helpers
from povu_b200 import api, synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE = os.path.join(ROOT, "oracle", "povu_oracle")
PREFIX = {"hairpin_inversion_subr": "ref", "LPA": "chm13", "multi_comp": "S001#1#chr1"}


def prefix_for(name):
    return PREFIX.get(name, "S001" if name in helpers.SYNTHETIC else "HG1")


def split(text, only_plain):
    lines = text.splitlines()
    hdr = [l for l in lines if l.startswith("#") and not l.startswith("##fileDate")]
    rec = [l for l in lines if not l.startswith("#")]
    if only_plain:      # the oracle restatement covers neither tangled regions nor the SNE pass
        rec = [l for l in rec if "VARTYPE=SUBR" not in l and "TANGLED=T" not in l]
    return hdr, rec


def gfa2vcf(pv, path, prefixes, chunk=100, via_pvst_dir=None):
    s = api.Session(pv, path)
    s.decompose()
    if via_pvst_dir is not None:        # decompose -> .pvst files -> call, like `povu decompose` + `povu call`
        s.write_pvst(via_pvst_dir)
        s.read_pvst(via_pvst_dir)
    s.call(prefixes)
    return s.vcf_text(chunk_size=chunk), s.call_stats(), s


def check_golden(pv, name, tmp):
    got, st, _ = gfa2vcf(pv, helpers.fixture_path(name, tmp), [prefix_for(name)])
    want = open(os.path.join(helpers.GOLDEN, "expected", name + ".vcf")).read()
    assert got == want, "byte-exact, including record order (LPA: all 2238 records, 1270 of them tangled; hairpin: the SUBR record)"
    assert st["hard_errors"] == 0
    return st


def oracle_vcf(gfa, prefixes, chunk=100):
    cmd = [ORACLE, "gfa2vcf", "-i", gfa, "-c", str(chunk)]
    for p in prefixes:
        cmd += ["-P", p]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-300:]
    return r.stdout


def check_vs_oracle(pv, g, prefixes, tmp, chunk=100, exact_order=True):
    gfa = os.path.join(tmp, "g.gfa")
    synth.write_gfa(g, gfa)
    got, st, _ = gfa2vcf(pv, gfa, prefixes, chunk)
    want = oracle_vcf(gfa, prefixes, chunk)
    gh, gr = split(got, False)
    wh, wr = split(want, True)
    gr = [l for l in gr if "TANGLED=T" not in l and "VARTYPE=SUBR" not in l]      # the oracle restatement skips tangled regions and SNE
    assert gh == wh
    if exact_order:
        assert gr == wr
    else:
        assert sorted(gr) == sorted(wr)
    assert st["hard_errors"] == 0
    return st


REF = os.path.join(ROOT, "oracle", "_ref", "povu_ref")


def require_ref():
    """The compiled reference is the checker of the tangled / SNE / cyclic sub-paths: without it those tests FAIL (a skip
    would silently drop that parity).  It is built by oracle/Makefile (`__graft_entry__.build()`) and travels to the GPU box."""
    assert os.path.exists(REF), "oracle/_ref/povu_ref is missing: run `python -c 'import __graft_entry__ as g; g.build()'` where /root/reference exists"


def reference_vcf(gfa, prefixes, chunk=100):
    """VCF of the compiled, unmodified reference (oracle/_ref: built by oracle/Makefile, travels to the GPU box)."""
    cmd = [REF, "gfa2vcf", "-i", gfa, "-t", "1", "-c", str(chunk)]
    for p in prefixes:
        cmd += ["-P", p]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-300:]
    return r.stdout


def run_tangled_vs_reference(pv, tmp, n_seeds=24):
    """Tangled regions (untangle + lap alignment): VNTR-like repeat units that paths traverse 1..5 times.  The oracle
    restatement does not cover this sub-path, so the checker is the compiled reference itself."""
    require_ref()
    tangled = records = 0
    gfa = os.path.join(tmp, "t.gfa")
    for seed in range(n_seeds):
        g = synth.repeat_loops(seed, n_units=3 + seed % 5, n_paths=3 + seed % 6, max_laps=2 + seed % 4)
        synth.write_gfa(g, gfa)
        got, st, _ = gfa2vcf(pv, gfa, ["S001#1"], chunk=1 + seed % 7)
        want = reference_vcf(gfa, ["S001#1"], chunk=1 + seed % 7)
        gh, gr = split(got, False)
        wh, wr = split(want, False)
        assert gh == wh and gr == wr, f"seed {seed}"
        assert st["hard_errors"] == 0
        tangled += st["tangled"]; records += sum("TANGLED=T" in l for l in gr)
    assert tangled > n_seeds and records > n_seeds, (tangled, records)


def run_cyclic_vs_reference(pv, tmp, n_seeds=30):
    """Regions with a cycle inside: the BFS/Kahn sort fails and the vertex order comes from the haplotype laps
    (find_laps + gen_sort, graph.cpp:165-395).  Checker: the compiled reference; where it crashes (it reads
    v_ids[-1] / an empty optional on some inputs) the product must report hard errors instead of records."""
    require_ref()
    gfa = os.path.join(tmp, "c.gfa")
    compared = crashed = records = 0
    for seed in range(n_seeds):
        g = synth.cyclic_units(seed, n_units=2 + seed % 5, n_paths=3 + seed % 6, max_laps=1 + seed % 4)
        synth.write_gfa(g, gfa)
        cmd = [REF, "gfa2vcf", "-i", gfa, "-t", "1", "-P", "S001#1"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            crashed += 1
            try:
                _, st, _ = gfa2vcf(pv, gfa, ["S001#1"])
            except RuntimeError:
                continue
            assert st["hard_errors"] > 0, f"seed {seed}: the reference crashed, the product reported nothing"
            continue
        got, st, _ = gfa2vcf(pv, gfa, ["S001#1"])
        gh, gr = split(got, False)
        wh, wr = split(r.stdout, False)
        assert gh == wh and gr == wr, f"seed {seed}"
        assert st["hard_errors"] == 0 and st["sort_failed"] == 0
        compared += 1; records += len(gr)
    assert compared >= n_seeds // 2 and records > 20, (compared, crashed, records)


def run_inversions_vs_reference(pv, tmp, n_seeds=45):
    """SNE / SUBR pass: paths walked on the reverse strand turn regions into whole-region inversions (pin pairs ->
    chains -> extensions -> slice tree -> SUBR records), also inside tangled regions and with two ref paths.  Checker:
    the compiled reference.  Seeds where a merged slice length runs past the end of a walk are skipped: the reference
    reads beyond its arrays there (undefined behaviour), the product clamps and counts (`subr_clamped`)."""
    require_ref()
    gfa = os.path.join(tmp, "i.gfa")
    subr = pins = compared = clamped_compared = 0
    for seed in range(n_seeds):
        kind = seed % 3
        if kind == 0:
            g = synth.snp_indel_chain(8 + seed % 30, n_paths=3 + seed % 6, seed=seed)
        elif kind == 1:
            g = synth.nested_bubbles(900 + seed, n_top=4 + seed % 6, max_depth=1 + seed % 3, n_paths=3 + seed % 6, p_inv=0.1)
        else:
            g = synth.repeat_loops(seed, n_units=2 + seed % 4, n_paths=3 + seed % 5, max_laps=2 + seed % 3)
        n = len(g["paths"])
        rev = {i for i in range(1, n) if (seed * 7 + i * 3) % 4 < 2}
        if seed % 5 == 0:
            rev.add(0)
        synth.write_gfa(synth.reverse_paths(g, rev), gfa)
        pref = ["S001"] if seed % 4 == 0 else ["S001#1"]
        chunk = 1 + (seed * 5) % 40
        got, st, _ = gfa2vcf(pv, gfa, pref, chunk)
        assert st["hard_errors"] == 0
        pins += st["pin_pairs"]
        if st["subr_clamped"]:
            # a merged slice ran past the end of a walk: the reference reads beyond its arrays there (undefined behaviour), so
            # its SUBR lines cannot be compared -- every other record still must agree (unless the reference crashed outright)
            cmd = [REF, "gfa2vcf", "-i", gfa, "-t", "1", "-c", str(chunk)] + [x for p_ in pref for x in ("-P", p_)]
            r = subprocess.run(cmd, capture_output=True, text=True)
            if r.returncode == 0:
                nosubr = lambda t: [l for l in split(t, False)[1] if "VARTYPE=SUBR" not in l]
                assert nosubr(got) == nosubr(r.stdout), f"seed {seed} (clamped)"
                clamped_compared += 1
            continue
        want = reference_vcf(gfa, pref, chunk)
        assert split(got, False) == split(want, False), f"seed {seed}"
        compared += 1
        subr += sum("VARTYPE=SUBR" in l for l in got.splitlines())
    assert compared >= n_seeds * 0.7 and subr > n_seeds and pins > 2 * n_seeds, (compared, clamped_compared, subr, pins)


def run_all_synthetic(pv, tmp, n_seeds=12):
    seen_big = 0
    for s in range(n_seeds):
        g = synth.nested_bubbles(100 + s, n_top=6 + s % 5, max_depth=2 + s % 4, n_paths=2 + s % 7, shuffle_links=(s % 2 == 1))
        st = check_vs_oracle(pv, g, ["S001"], tmp, chunk=1 + (s * 7) % 50)
        seen_big += st["big_bfs"]
    for s in range(4):
        g = synth.snp_indel_chain(150 + 40 * s, n_paths=4 + 2 * s, seed=20 + s)
        check_vs_oracle(pv, g, ["S001"], tmp)
    # two reference paths at once (-P S001 selects both haplotypes of the sample)
    g = synth.snp_indel_chain(80, n_paths=6, seed=5)
    check_vs_oracle(pv, g, ["S001", "S002"], tmp)
    # several components, paths per component
    g = synth.concat_components([synth.nested_bubbles(41, n_top=5), synth.snp_indel_chain(30, 4, seed=6), synth.nested_bubbles(43, n_top=7, max_depth=4)])
    check_vs_oracle(pv, g, ["S001#1#chr1"], tmp)
    check_vs_oracle(pv, g, ["S001"], tmp)
    return seen_big


def run_deep_regions(pv, tmp):
    """Few paths over deep nesting: parents are selected with large regions, so the global-arena BFS tiers,
    the (region, vertex) hash and multi-chunk processing (tiny cell budget) are all exercised."""
    os.environ["POVU_B200_CELL_BUDGET"] = "64"
    try:
        big = 0; mx = 0
        for s in range(6):
            g = synth.nested_bubbles(300 + s, n_top=4, max_depth=6, n_paths=2 + s % 2, p_multi=0.6)
            st = check_vs_oracle(pv, g, ["S001"], tmp)
            big += st["big_bfs"]; mx = max(mx, st["max_region"])
            assert st["chunks"] >= 1
        assert big > 0 and mx > 16, (big, mx)
    finally:
        del os.environ["POVU_B200_CELL_BUDGET"]


def check_pvst_roundtrip(pv, name, tmp):
    path = helpers.fixture_path(name, tmp)
    a, _, s = gfa2vcf(pv, path, [prefix_for(name)])
    d = tempfile.mkdtemp(dir=tmp)
    b, _, s2 = gfa2vcf(pv, path, [prefix_for(name)], via_pvst_dir=d)
    assert a == b
    files = {int(f[:-5]): open(os.path.join(d, f)).read() for f in os.listdir(d)}
    assert files == helpers.expected_pvsts(name) == s.pvst_texts()


# ---------------------------------------------------------------------------------------------------------------------
# perturbed graphs through the CLI: nested bubbles + random chords (any sides: inversions, back links, self-loops) and
# haplotypes that are random walks over the result -- regions the generators above never build (tangled regions inside
# nested ones, walks that enter a region in its middle, paths that stop inside a region)
# ---------------------------------------------------------------------------------------------------------------------
def random_walks(g, rng, n_paths, max_len):
    V = len(g["ids"])
    adj = {}
    for a, sa, b, sb in zip(g["v1"].tolist(), g["s1"].tolist(), g["v2"].tolist(), g["s2"].tolist()):
        adj.setdefault((a, sa), []).append((b, sb)); adj.setdefault((b, sb), []).append((a, sa))
    out = []
    for _ in range(n_paths):
        v = int(rng.integers(0, V)); strand = int(rng.integers(0, 2))      # strand 0 = forward: in by the left side, out by the right
        st, sr = [v], [strand]
        for _ in range(int(rng.integers(2, max_len))):
            nb = adj.get((v, 1 if strand == 0 else 0))
            if not nb:
                break
            v, ws = nb[int(rng.integers(0, len(nb)))]
            strand = 0 if ws == 0 else 1
            st.append(v); sr.append(strand)
        out.append((np.array(st, dtype=np.uint32), np.array(sr, dtype=np.uint8)))
    return out


def perturbed_graph(seed):
    rng = np.random.default_rng(10_000 + seed)
    g = synth.nested_bubbles(seed, n_top=4 + seed % 5, max_depth=1 + seed % 4, n_paths=6)
    V = len(g["ids"])
    extra = [(int(rng.integers(0, V)), int(rng.integers(0, 2)), int(rng.integers(0, V)), int(rng.integers(0, 2))) for _ in range(int(rng.integers(0, 4)))]
    if extra:
        L = np.array(extra, dtype=np.int64)
        g["v1"] = np.concatenate([g["v1"], L[:, 0].astype(np.uint32)]); g["s1"] = np.concatenate([g["s1"], L[:, 1].astype(np.uint8)])
        g["v2"] = np.concatenate([g["v2"], L[:, 2].astype(np.uint32)]); g["s2"] = np.concatenate([g["s2"], L[:, 3].astype(np.uint8)])
    paths = list(g["paths"][:4])
    for i, (st, sr) in enumerate(random_walks(g, rng, 4, 40)):
        paths.append((f"S{(4 + i) // 2 + 1:03d}#{(4 + i) % 2 + 1}#chr1", st, sr))
    g["paths"] = paths
    return g


def perturbed_graph_big(seed):
    """the same idea at 3-5 times the size: 10-24 top-level bubbles nested up to 5 deep, 2-11 chords (70 % of them between
    vertices at most 12 indices apart: local structural variation), 5 template walks + 5 random walks of up to 80 steps"""
    rng = np.random.default_rng(50_000 + seed)
    g = synth.nested_bubbles(seed, n_top=10 + seed % 15, max_depth=2 + seed % 4, n_paths=8)
    V = len(g["ids"])
    ex = []
    for _ in range(int(rng.integers(2, 12))):
        a = int(rng.integers(0, V))
        b = int(np.clip(a + int(rng.integers(-12, 13)), 0, V - 1)) if rng.uniform() < 0.7 else int(rng.integers(0, V))
        ex.append((a, int(rng.integers(0, 2)), b, int(rng.integers(0, 2))))
    L = np.array(ex, dtype=np.int64)
    g["v1"] = np.concatenate([g["v1"], L[:, 0].astype(np.uint32)]); g["s1"] = np.concatenate([g["s1"], L[:, 1].astype(np.uint8)])
    g["v2"] = np.concatenate([g["v2"], L[:, 2].astype(np.uint32)]); g["s2"] = np.concatenate([g["s2"], L[:, 3].astype(np.uint8)])
    paths = list(g["paths"][:5])
    for i, (st, sr) in enumerate(random_walks(g, rng, 5, 80)):
        paths.append((f"S{(5 + i) // 2 + 1:03d}#{(5 + i) % 2 + 1}#chr1", st, sr))
    g["paths"] = paths
    return g


def check_perturbed_cli(cli, tmp, seeds, env=None, build=None, args=None):
    """gfa2vcf of `cli` vs the compiled reference, whole output byte for byte (order included).  Seeds on which the reference
    itself crashes are skipped; where the product reports a clamped SUBR slice (the reference reads past the end of a walk
    there: undefined) every other line must still agree.  Returns (graphs compared, records, tangled, SUBR)."""
    require_ref()
    n = nrec = ntang = nsubr = 0
    strip = lambda t: [l for l in t.splitlines() if not l.startswith("##fileDate")]
    for seed in seeds:
        g = (build or perturbed_graph)(seed)
        gfa = os.path.join(tmp, f"pert{seed}.gfa")
        synth.write_gfa(g, gfa)
        sel = args or ["-P", g["paths"][0][0].split("#")[0]]
        r = subprocess.run([REF, "gfa2vcf", "-i", gfa, "-t", "1", *sel], capture_output=True, text=True)
        if r.returncode != 0:
            continue
        m = subprocess.run([cli, "gfa2vcf", "-i", gfa, "-t", "1", *sel], capture_output=True, text=True, env=dict(os.environ, **(env or {})))
        if m.returncode != 0 and "the reference aborts on this input" in m.stderr:
            continue      # gen_cxt takes front() of an empty vector there (overlay.cpp:307-325): the reference survives by luck, the product refuses
        assert m.returncode == 0, (seed, m.stderr[-300:])
        a, b = strip(r.stdout), strip(m.stdout)
        if "SUBR records had a merged length past the end of a walk" in m.stderr:
            a = [l for l in a if "VARTYPE=SUBR" not in l and "\t60\tPASS\t" in l or l.startswith("#")]
            a = [l for l in a if l.startswith("#") or "VARTYPE=" in l]      # (the reference's over-read line has no intact INFO column)
            b = [l for l in b if "VARTYPE=SUBR" not in l]
        assert a == b, (seed, [l[:120] for l in a if l not in set(b)][:2], [l[:120] for l in b if l not in set(a)][:2])
        n += 1; nrec += sum(1 for l in a if not l.startswith("#")); ntang += sum("TANGLED=T" in l for l in a); nsubr += sum("VARTYPE=SUBR" in l for l in a)
    return n, nrec, ntang, nsubr


def check_perturbed_api(pv, tmp, seeds, build=None):
    """the same comparison through the in-process operators (one CUDA context for all graphs instead of one per CLI run):
    Session.from_graph -> decompose -> call -> VCF text vs the compiled reference's stdout"""
    require_ref()
    n = nrec = 0
    strip = lambda t: [l for l in t.splitlines() if not l.startswith("##fileDate")]
    for seed in seeds:
        g = (build or perturbed_graph)(seed)
        gfa = os.path.join(tmp, f"perta{seed}.gfa")
        synth.write_gfa(g, gfa)
        prefix = g["paths"][0][0].split("#")[0]
        r = subprocess.run([REF, "gfa2vcf", "-i", gfa, "-t", "1", "-P", prefix], capture_output=True, text=True)
        if r.returncode != 0:
            continue
        s = api.Session.from_graph(pv, g)
        try:
            s.decompose(); s.call([prefix])
            st = s.call_stats()
            if st["hard_errors"]:
                continue      # the product refuses (CLI: "the reference aborts on this input")
            a, b = strip(r.stdout), strip(s.vcf_text(with_date=False))
            if st["subr_clamped"]:
                a = [l for l in a if l.startswith("#") or ("VARTYPE=" in l and "VARTYPE=SUBR" not in l)]
                b = [l for l in b if "VARTYPE=SUBR" not in l]
            assert a == b, (seed, [l[:120] for l in a if l not in set(b)][:2], [l[:120] for l in b if l not in set(a)][:2])
        finally:
            s.close()
        n += 1; nrec += sum(1 for l in a if not l.startswith("#"))
    return n, nrec


def check_perturbed_files_and_restrict(cli, tmp, seeds):
    """decompose -> .pvst files (byte-identical to the reference's) -> `call -f dir` on the reference's files, every third
    graph with a `-g ref:start-end` window over the middle of the reference walk"""
    import shutil
    require_ref()
    n = nrec = 0
    strip = lambda t: [l for l in t.splitlines() if not l.startswith("##fileDate")]
    for seed in seeds:
        g = perturbed_graph_big(seed) if seed % 2 else perturbed_graph(seed)
        gfa = os.path.join(tmp, f"pf{seed}.gfa")
        synth.write_gfa(g, gfa)
        fr, fm = os.path.join(tmp, f"pf{seed}.r"), os.path.join(tmp, f"pf{seed}.m")
        for d in (fr, fm):
            shutil.rmtree(d, ignore_errors=True); os.makedirs(d)
        ref0 = g["paths"][0][0]
        length = sum(len(g["seqs"][v]) for v in g["paths"][0][1].tolist())
        if subprocess.run([REF, "decompose", "-i", gfa, "-o", fr, "-t", "1"], capture_output=True).returncode != 0:
            continue
        assert subprocess.run([cli, "decompose", "-i", gfa, "-o", fm], capture_output=True).returncode == 0, seed
        assert {f: open(os.path.join(fr, f)).read() for f in os.listdir(fr)} == {f: open(os.path.join(fm, f)).read() for f in os.listdir(fm)}, (seed, "pvst")
        sel = ["-P", ref0.split("#")[0]] + (["-g", f"{ref0}:{int(length * 0.2)}-{int(length * 0.7)}"] if seed % 3 == 0 else [])
        r = subprocess.run([REF, "call", "-i", gfa, "-f", fr, "-t", "1", *sel], capture_output=True, text=True)
        if r.returncode != 0:
            continue
        m = subprocess.run([cli, "call", "-i", gfa, "-f", fr, "-t", "1", *sel], capture_output=True, text=True)
        if m.returncode != 0 and "the reference aborts on this input" in m.stderr:
            continue
        assert m.returncode == 0, (seed, m.stderr[-300:])
        a, b = strip(r.stdout), strip(m.stdout)
        if "SUBR records had a merged length past the end of a walk" in m.stderr:
            a = [l for l in a if l.startswith("#") or ("VARTYPE=" in l and "VARTYPE=SUBR" not in l)]
            b = [l for l in b if "VARTYPE=SUBR" not in l]
        assert a == b, (seed, sel)
        n += 1; nrec += sum(1 for l in a if not l.startswith("#"))
    return n, nrec


def hub_graph(seed):
    """nested bubbles plus 1-3 hubs (one vertex side linked to 3-8 random vertex sides) and a few more chords: DFS-tree vertices
    with many children whose brackets reach different heights -- where the reference's choice of capping edge matters most
    (about three graphs in four take the re-derived classes of pv_refnum.cuh).  No haplotypes: a decompose family."""
    rng = np.random.default_rng(90_000 + seed)
    g = synth.nested_bubbles(seed, n_top=6 + seed % 10, max_depth=1 + seed % 4, n_paths=2)
    V = len(g["ids"])
    ex = []
    for _ in range(1 + seed % 3):
        h, hs = int(rng.integers(0, V)), int(rng.integers(0, 2))
        for _ in range(int(rng.integers(3, 9))):
            ex.append((h, hs, int(rng.integers(0, V)), int(rng.integers(0, 2))))
    for _ in range(int(rng.integers(0, 5))):
        ex.append((int(rng.integers(0, V)), int(rng.integers(0, 2)), int(rng.integers(0, V)), int(rng.integers(0, 2))))
    L = np.array(ex, dtype=np.int64)
    g["v1"] = np.concatenate([g["v1"], L[:, 0].astype(np.uint32)]); g["s1"] = np.concatenate([g["s1"], L[:, 1].astype(np.uint8)])
    g["v2"] = np.concatenate([g["v2"], L[:, 2].astype(np.uint32)]); g["s2"] = np.concatenate([g["s2"], L[:, 3].astype(np.uint8)])
    return g


def check_hub_graphs(pv, tmp, seeds, ref_every=4):
    """decompose vs the oracle on every graph and vs the compiled reference's .pvst on every `ref_every`-th"""
    import oracle_binding as ob
    require_ref()
    n = faithful = nref = 0
    for seed in seeds:
        g = hub_graph(seed)
        ep1, ep2 = api.edge_ports(g)
        f = pv.decompose(len(g["ids"]), ep1, ep2)
        assert ob.flubble_rows(f, g["ids"]) == ob.oracle_decompose(g), seed
        n += 1; faithful += 1 if f.faithful_classes else 0
        if seed % ref_every == 0:
            gfa = os.path.join(tmp, f"hub{seed}.gfa"); d = os.path.join(tmp, f"hub{seed}.d")
            os.makedirs(d, exist_ok=True)
            synth.write_gfa(g, gfa)
            if subprocess.run([REF, "decompose", "-i", gfa, "-o", d, "-t", "1"], capture_output=True).returncode != 0:
                continue
            assert f.pvst_texts(g["ids"]) == {int(x[:-5]): open(os.path.join(d, x)).read() for x in os.listdir(d)}, seed
            nref += 1
    return n, faithful, nref


def check_subflubble_forests(cli, tmp, seeds, with_lpa=True):
    """`call -f` on forests written by the REFERENCE's `decompose -s`: tiny (T) and parallel (O) vertices are flubble-like and are
    called, the concealed / smothered / midi clan and everything below it is not (pvst.hpp:110-125, color.cpp:117-125), and
    vertices that no children list reaches from the root -- the sub-flubble passes leave some -- are never looked at"""
    import shutil
    require_ref()
    n = nrec = 0
    fams = {}
    strip = lambda t: [l for l in t.splitlines() if not l.startswith("##fileDate")]
    cases = []
    for seed in seeds:
        g = perturbed_graph_big(seed) if seed % 2 else perturbed_graph(seed)
        gfa = os.path.join(tmp, f"sf{seed}.gfa")
        synth.write_gfa(g, gfa)
        cases.append((gfa, g["paths"][0][0].split("#")[0]))
    if with_lpa:
        cases.append((helpers.fixture_path("LPA", tmp), "chm13"))
    for gfa, prefix in cases:
        d = gfa + ".forest"
        shutil.rmtree(d, ignore_errors=True); os.makedirs(d)
        if subprocess.run([REF, "decompose", "-i", gfa, "-o", d, "-t", "1", "-s"], capture_output=True).returncode != 0:
            continue
        for f in os.listdir(d):
            for l in open(os.path.join(d, f)):
                fams[l[0]] = fams.get(l[0], 0) + 1
        r = subprocess.run([REF, "call", "-i", gfa, "-f", d, "-t", "1", "-P", prefix], capture_output=True, text=True)
        if r.returncode != 0:
            continue
        m = subprocess.run([cli, "call", "-i", gfa, "-f", d, "-t", "1", "-P", prefix], capture_output=True, text=True)
        if m.returncode != 0 and "the reference aborts on this input" in m.stderr:
            continue
        assert m.returncode == 0, (gfa, m.stderr[-300:])
        a, b = strip(r.stdout), strip(m.stdout)
        if "SUBR records had a merged length past the end of a walk" in m.stderr:
            a = [l for l in a if l.startswith("#") or ("VARTYPE=" in l and "VARTYPE=SUBR" not in l)]
            b = [l for l in b if "VARTYPE=SUBR" not in l]
        assert a == b, gfa
        n += 1; nrec += sum(1 for l in a if not l.startswith("#"))
    return n, nrec, fams
