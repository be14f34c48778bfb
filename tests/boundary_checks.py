"""Drop-in boundary checks shared by the CPU (host rehearsal) and GPU test files: the `povu` CLI contract
(/root/reference/app/cli/cli.cpp, tests/integration_tests/gfa2vcf_tests.cc:113-136) and the povu_ffi.h C API
(/root/reference/povu-rs/tests/{builder,graph}_tests.rs expectations)."""
import ctypes as C
import os
import re
import subprocess
import threading

import helpers

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def run(cli, *args, env=None):
    return subprocess.run([cli, *args], capture_output=True, text=True, env=dict(os.environ, **(env or {})))


def nodate(t):
    return "".join(l for l in t.splitlines(True) if not l.startswith("##fileDate"))


def golden(name):
    return open(os.path.join(helpers.GOLDEN, "expected", name + ".vcf")).read()


def check_cli(cli, tmp):
    fx = helpers.fixture_path
    # gfa2vcf to stdout, byte-exact against the compiled reference (minus the date line)
    for name in ["minimal_substitution", "nested_deletion", "two_ordered_substitutions", "complex_substitution_span"]:
        r = run(cli, "gfa2vcf", "-i", fx(name), "-t", "1", "-P", "HG1")
        assert r.returncode == 0, r.stderr
        assert re.search(r"^##fileDate=\d{8}$", r.stdout, re.M)
        assert nodate(r.stdout) == golden(name)
    # positional refs and --path-prefix= spellings select the same thing
    a = run(cli, "gfa2vcf", "-i", fx("insertion_flubble"), "HG1").stdout
    b = run(cli, "gfa2vcf", "--input-gfa", fx("insertion_flubble"), "--path-prefix=HG1").stdout
    assert a == b and nodate(a) == golden("insertion_flubble")
    # decompose -> .pvst files, then call -f dir: same VCF, files identical to the reference's
    d = os.path.join(tmp, "forest"); os.mkdir(d)
    r = run(cli, "decompose", "-i", fx("multi_comp"), "-o", d)
    assert r.returncode == 0, r.stderr
    files = {int(f[:-5]): open(os.path.join(d, f)).read() for f in os.listdir(d)}
    assert files == helpers.expected_pvsts("multi_comp")
    r = run(cli, "call", "-i", fx("multi_comp"), "-f", d, "-P", "S001#1#chr1")
    assert r.returncode == 0 and nodate(r.stdout) == golden("multi_comp")
    # POVU_STAGE_COST_TRACE=1: the reference's six stage lines (stage_cost.cpp:59-94), same names / order / fields; the item
    # counts it defines on the inputs agree with the compiled reference where one exists
    d2 = os.path.join(tmp, "forest2"); os.mkdir(d2)
    r = run(cli, "decompose", "-i", fx("multi_comp"), "-o", d2, env={"POVU_STAGE_COST_TRACE": "1"})
    assert r.returncode == 0
    lines = [l for l in r.stderr.splitlines() if l.startswith("povu-stage-cost")]
    assert lines[0] == "povu-stage-cost-trace schema=povu.stage-cost.v1"
    names = [re.match(r"povu-stage-cost contract=(\w+) trace=([\w-]+) calls=(\d+) input_items=(\d+) output_items=(\d+) elapsed_ns=(\d+)$", l).groups() for l in lines[1:]]
    assert [(n[0], n[1]) for n in names] == [("componentDecomposition", "component-decomposition"), ("tipDummyAugmentation", "tip-dummy-augmentation"),
                                              ("traversalFrameConstruction", "traversal-frame-construction"), ("cycleClassAssignment", "cycle-class-assignment"),
                                              ("boundaryEmission", "flubble-boundary-extraction"), ("hierarchyConstruction", "flubble-hierarchy-construction")]
    import call_checks as cc0
    if os.path.exists(cc0.REF):
        rr = subprocess.run([cc0.REF, "decompose", "-i", fx("multi_comp"), "-o", d2, "-t", "1"], capture_output=True, text=True, env=dict(os.environ, POVU_STAGE_COST_TRACE="1"))
        ref = {m[0]: m for m in re.findall(r"povu-stage-cost contract=(\w+) trace=[\w-]+ calls=(\d+) input_items=(\d+) output_items=(\d+)", rr.stderr)}
        got = {n[0]: n for n in names}
        assert got["componentDecomposition"][3:5] == ref["componentDecomposition"][2:4]                 # V + E in, components out
        assert got["boundaryEmission"][4] == ref["boundaryEmission"][3]                                  # flubble boundaries found
        assert got["traversalFrameConstruction"][4] == ref["traversalFrameConstruction"][3]              # class-stack entries
    # -o dir: one <prefix>.vcf per prefix (to_vcf.hpp:83-101); -r file
    out = os.path.join(tmp, "vcfs")
    pf = os.path.join(tmp, "refs.txt"); open(pf, "w").write("S001\n")
    r = run(cli, "gfa2vcf", "-i", fx("chain_s3"), "-r", pf, "-o", out)
    assert r.returncode == 0 and r.stdout == ""
    assert nodate(open(os.path.join(out, "S001.vcf")).read()) == golden("chain_s3")
    # chunk size changes record grouping exactly like the reference (genomics.cpp:114-156): compare with the oracle
    import call_checks as cc
    r = run(cli, "gfa2vcf", "-i", fx("chain_s3"), "-P", "S001", "-c", "7")
    assert cc.split(r.stdout, False)[1] == cc.split(cc.oracle_vcf(fx("chain_s3"), ["S001"], 7), True)[1]
    # rejection: exit != 0, no signal, stderr names the problem (gfa2vcf_tests.cc:113-136)
    for name in helpers.NEGATIVE:
        r = run(cli, "gfa2vcf", "-i", fx(name), "-P", "HG1")
        assert r.returncode == 1 and "Invalid GFA" in r.stderr and r.stdout == ""
    z = os.path.join(tmp, "z.gfa"); open(z, "w").write("H\tVN:Z:1.0\nS\t1\tA\nZ\tfoo\n")
    r = run(cli, "gfa2vcf", "-i", z, "-P", "x")
    assert r.returncode == 1 and "[povu::subcommands::do_gfa2vcf] Error: Invalid GFA" in r.stderr and "unsupported record type 'Z' on line 3" in r.stderr
    s = os.path.join(tmp, "s.gfa"); open(s, "w").write("S\t1\n")
    r = run(cli, "gfa2vcf", "-i", s, "-P", "x")
    assert r.returncode == 1 and "S record on line 1 is missing a sequence" in r.stderr
    # CLI parse errors: message + usage on stderr, exit 0, nothing done (cli.cpp:355-368, main.cpp:15)
    r = run(cli, "gfa2vcf", "-i", fx("minimal_substitution"))
    assert r.returncode == 0 and r.stdout == "" and "povu {OPTIONS}" in r.stderr
    r = run(cli, "gfa2vcf", "-P", "HG1")
    assert r.returncode == 0 and r.stdout == "" and "required" in r.stderr
    r = run(cli, "--version")
    assert r.returncode == 0 and r.stdout.strip() == "0.0.1-alpha"
    # missing forest
    e = os.path.join(tmp, "empty"); os.mkdir(e)
    r = run(cli, "call", "-i", fx("minimal_substitution"), "-f", e, "-P", "HG1")
    assert r.returncode == 1 and "Could not find pvst files" in r.stderr
    # call -g ref:start-end (rov.cpp:51-183): same records / header / exit code as the compiled reference
    import call_checks as cc2
    if os.path.exists(cc2.REF):
        from povu_b200 import synth
        gfa = os.path.join(tmp, "restrict.gfa")
        synth.write_gfa(synth.snp_indel_chain(300, n_paths=6, seed=3, with_seqs=True), gfa)
        fd = os.path.join(tmp, "rforest"); os.mkdir(fd)
        assert run(cli, "decompose", "-i", gfa, "-o", fd).returncode == 0
        for reg in ["S001#1#chr1:100-400", "S002#2#chr1:1-50", "nope:1-5", "S001#1#chr1:5-5", "S001#1#chr1:a-5", "S001#1#chr15",
                    "S001#1#chr1:-5", "S001#1#chr1:10-4000000000000", "S001#1#chr1:1-99999999999999999999999"]:
            want = run(cc2.REF, "call", "-i", gfa, "-f", fd, "-P", "S001#1", "-g", reg)
            got = run(cli, "call", "-i", gfa, "-f", fd, "-P", "S001#1", "-g", reg)
            assert got.returncode == want.returncode and nodate(got.stdout) == nodate(want.stdout), reg
            errs = [l[l.index("]") + 1:].split("m", 1)[-1].strip() for l in want.stderr.splitlines() if "ERR" in l]
            for e in errs:
                assert e in got.stderr, (reg, e)
    # outside the hot path: loud failure, never a silent no-op
    r = run(cli, "gfa2vcf", "-i", fx("minimal_substitution"), "-P", "HG1", "-s")
    assert r.returncode == 1 and "subflubbles" in r.stderr


class PovuError(C.Structure):
    _fields_ = [("code", C.c_int), ("message", C.c_char_p)]


class PovuStep(C.Structure):
    _fields_ = [("vertex_id", C.c_uint64), ("orientation", C.c_int)]


class PovuEdge(C.Structure):
    _fields_ = [("from_id", C.c_uint64), ("from_orientation", C.c_int), ("to_id", C.c_uint64), ("to_orientation", C.c_int)]


class PovuVertex(C.Structure):
    _fields_ = [("id", C.c_uint64), ("sequence", C.c_char_p), ("sequence_len", C.c_size_t)]


class PovuFlubble(C.Structure):
    _fields_ = [("id", C.c_uint64), ("type_name", C.c_char_p), ("start_vertex_id", C.c_uint64), ("end_vertex_id", C.c_uint64),
                ("walks", C.c_void_p), ("walk_lengths", C.c_void_p), ("walks_count", C.c_size_t)]


def ffi(lib_path):
    L = C.CDLL(lib_path)
    vp, sz, u64 = C.c_void_p, C.c_size_t, C.c_uint64
    EP = C.POINTER(PovuError)
    sig = {
        "povu_graph_new": (vp, [sz, sz, sz]), "povu_graph_from_gfa": (vp, [C.c_char_p, EP]), "povu_graph_free": (None, [vp]),
        "povu_graph_add_vertex": (sz, [vp, u64, C.c_char_p]), "povu_graph_add_edge": (sz, [vp, u64, C.c_int, u64, C.c_int]),
        "povu_graph_add_path": (C.c_bool, [vp, C.c_char_p, C.POINTER(PovuStep), sz]), "povu_graph_finalize": (None, [vp]),
        "povu_graph_vertex_count": (sz, [vp]), "povu_graph_edge_count": (sz, [vp]), "povu_graph_path_count": (sz, [vp]),
        "povu_graph_get_vertices": (C.POINTER(PovuVertex), [vp, C.POINTER(sz)]), "povu_graph_get_edges": (C.POINTER(PovuEdge), [vp, C.POINTER(sz)]),
        "povu_vertices_free": (None, [C.POINTER(PovuVertex), sz]), "povu_edges_free": (None, [C.POINTER(PovuEdge), sz]),
        "povu_graph_set_references_from_prefixes": (C.c_bool, [vp, C.POINTER(C.c_char_p), sz, EP]),
        "povu_graph_set_references_from_file": (C.c_bool, [vp, C.c_char_p, EP]),
        "povu_graph_find_flubbles": (vp, [vp, EP]), "povu_flubbles_free": (None, [vp]), "povu_flubbles_count": (sz, [vp]),
        "povu_flubbles_get": (C.POINTER(PovuFlubble), [vp, sz]), "povu_flubble_free": (None, [C.POINTER(PovuFlubble)]),
        "povu_flubbles_get_pvst_tree": (vp, [vp]), "povu_pvst_tree_free": (None, [vp]), "povu_pvst_tree_vertex_count": (sz, [vp]),
        "povu_flubbles_call_variants": (vp, [vp, EP]), "povu_vcf_to_string": (vp, [vp, C.POINTER(sz)]), "povu_vcf_write_to_file": (C.c_bool, [vp, C.c_char_p, EP]),
        "povu_vcf_free": (None, [vp]), "povu_string_free": (None, [vp]), "povu_gfa_to_vcf": (C.c_bool, [C.c_char_p, C.c_char_p, C.c_char_p, EP]),
        "povu_error_free": (None, [EP]),
    }
    for n, (r, a) in sig.items():
        f = getattr(L, n); f.restype = r; f.argtypes = a
    return L


def check_ffi_builder(L):
    """builder_tests.rs expectations that need no device."""
    g = L.povu_graph_new(4, 4, 1)
    assert g
    assert L.povu_graph_add_vertex(g, 1, b"ACGT") == 0 and L.povu_graph_add_vertex(g, 2, b"G") == 1 and L.povu_graph_add_vertex(g, 3, b"TT") == 2
    assert L.povu_graph_add_edge(g, 1, 1, 2, 0) == 0           # 1.right -- 2.left
    assert L.povu_graph_add_edge(g, 2, 1, 2, 0) == 1           # self-loop allowed
    assert L.povu_graph_add_edge(g, 2, 1, 3, 0) == 2
    assert L.povu_graph_add_edge(g, 9, 0, 1, 0) == C.c_size_t(-1).value   # unknown vertex
    assert L.povu_graph_add_vertex(g, 7, None) == C.c_size_t(-1).value
    assert L.povu_graph_add_vertex(g, 2, b"C") == C.c_size_t(-1).value          # duplicate id
    L.povu_graph_finalize(g)
    assert (L.povu_graph_vertex_count(g), L.povu_graph_edge_count(g), L.povu_graph_path_count(g)) == (3, 3, 0)
    n = C.c_size_t()
    e = L.povu_graph_get_edges(g, C.byref(n))
    assert n.value == 3 and (e[0].from_id, e[0].from_orientation, e[0].to_id, e[0].to_orientation) == (1, 1, 2, 0)   # orientation round trip (builder_tests.rs:171-199)
    L.povu_edges_free(e, n)
    v = L.povu_graph_get_vertices(g, C.byref(n))
    assert n.value == 3 and v[0].sequence == b"ACGT" and v[0].sequence_len == 4 and v[2].id == 3
    L.povu_vertices_free(v, n)
    L.povu_graph_free(g); L.povu_graph_free(None); L.povu_flubbles_free(None); L.povu_vcf_free(None)
    err = PovuError()
    assert not L.povu_graph_from_gfa(b"/nonexistent/x.gfa", C.byref(err)) and err.code == 1 and b"does not exist" in err.message
    L.povu_error_free(C.byref(err))
    assert not L.povu_graph_from_gfa(b"", C.byref(err)) and b"must not be empty" in err.message
    L.povu_error_free(C.byref(err))
    bad = helpers.fixture_path("unsupported_overlap")
    assert not L.povu_graph_from_gfa(bad.encode(), C.byref(err)) and b"Invalid GFA" in err.message
    L.povu_error_free(C.byref(err))


def _vcf_via_ffi(L, g, prefixes):
    err = PovuError()
    arr = (C.c_char_p * len(prefixes))(*[p.encode() for p in prefixes])
    assert L.povu_graph_set_references_from_prefixes(g, arr, len(prefixes), C.byref(err))
    f = L.povu_graph_find_flubbles(g, C.byref(err))
    assert f, err.message
    o = L.povu_flubbles_call_variants(f, C.byref(err))
    assert o, err.message
    n = C.c_size_t()
    p = L.povu_vcf_to_string(o, C.byref(n))
    text = C.string_at(p, n.value).decode()
    L.povu_string_free(p); L.povu_vcf_free(o)
    return f, text


def check_ffi_compute(L, tmp):
    """The entry points the reference leaves as stubs, made real."""
    name = "nested_s1"
    g = L.povu_graph_from_gfa(helpers.fixture_path(name).encode(), None)
    assert g
    # a graph loaded from GFA keeps its links / walks / sequences in HBM (device parser): the getters bring them to the host
    src = helpers.read_gfa(helpers.fixture_path(name))
    n = C.c_size_t()
    v = L.povu_graph_get_vertices(g, C.byref(n))
    assert n.value == len(src["ids"]) and all(v[i].id == int(src["ids"][i]) and v[i].sequence == src["seqs"][i].encode() for i in range(n.value))
    L.povu_vertices_free(v, n)
    e = L.povu_graph_get_edges(g, C.byref(n))
    assert n.value == len(src["v1"])
    for i in range(n.value):       # PovuOrientation names the vertex END for both endpoints (povu_ffi.cpp:148-156): 0 = left, 1 = right
        assert (e[i].from_id, e[i].from_orientation, e[i].to_id, e[i].to_orientation) == (int(src["ids"][src["v1"][i]]), int(src["s1"][i]), int(src["ids"][src["v2"][i]]), int(src["s2"][i])), i
    L.povu_edges_free(e, n)
    assert L.povu_graph_path_count(g) == len(src["paths"])
    f, text = _vcf_via_ffi(L, g, ["S001"])
    assert nodate(text) == golden(name)
    nfl = L.povu_flubbles_count(f)
    want = helpers.expected_pvsts(name)[1]
    assert nfl == len(want.splitlines()) - 2
    t = L.povu_flubbles_get_pvst_tree(f)
    assert L.povu_pvst_tree_vertex_count(t) == nfl + 1
    L.povu_pvst_tree_free(t)
    fl = L.povu_flubbles_get(f, 0)
    lab = want.splitlines()[2].split("\t")[2]
    a, z = re.findall(r"\d+", lab)
    assert fl.contents.type_name == b"flubble" and {fl.contents.start_vertex_id, fl.contents.end_vertex_id} == {int(a), int(z)}
    L.povu_flubble_free(fl)
    assert not L.povu_flubbles_get(f, nfl)
    L.povu_flubbles_free(f); L.povu_graph_free(g)
    # the same graph through the builder API gives the same VCF
    src = helpers.read_gfa(helpers.fixture_path("minimal_substitution"))
    b = L.povu_graph_new(0, 0, 0)
    for i, s in zip(src["ids"], src["seqs"]):
        L.povu_graph_add_vertex(b, int(i), s.encode())
    for v1, s1, v2, s2 in zip(src["v1"], src["s1"], src["v2"], src["s2"]):
        assert L.povu_graph_add_edge(b, int(src["ids"][v1]), int(s1), int(src["ids"][v2]), int(s2)) != C.c_size_t(-1).value
    for nm, steps, strands in src["paths"]:
        arr = (PovuStep * len(steps))(*[PovuStep(int(src["ids"][v]), int(o)) for v, o in zip(steps, strands)])
        assert L.povu_graph_add_path(b, nm.encode(), arr, len(steps))
    L.povu_graph_finalize(b)
    assert L.povu_graph_path_count(b) == len(src["paths"])
    f, text = _vcf_via_ffi(L, b, ["HG1"])
    assert nodate(text) == golden("minimal_substitution")
    nfl1 = L.povu_flubbles_count(f)
    L.povu_flubbles_free(f)
    # mutating a graph after a call: a second component is added, every table derived from the old graph must be rebuilt
    # (the text tables / forest / labels of the previous graph used to survive: out-of-bounds reads in the text kernels)
    V0 = max(int(i) for i in src["ids"]) + 1
    for i, sq in zip(src["ids"], src["seqs"]):
        assert L.povu_graph_add_vertex(b, V0 + int(i), sq.encode()) != C.c_size_t(-1).value
    for v1, s1, v2, s2 in zip(src["v1"], src["s1"], src["v2"], src["s2"]):
        assert L.povu_graph_add_edge(b, V0 + int(src["ids"][v1]), int(s1), V0 + int(src["ids"][v2]), int(s2)) != C.c_size_t(-1).value
    for nm, steps, strands in src["paths"]:
        arr = (PovuStep * len(steps))(*[PovuStep(V0 + int(src["ids"][v]), int(o)) for v, o in zip(steps, strands)])
        assert L.povu_graph_add_path(b, ("ZZ" + nm[2:]).replace("chr1", "chr2").encode(), arr, len(steps))
    f2, text2 = _vcf_via_ffi(L, b, ["HG1"])
    body = lambda t: [l.split("\t")[3:5] for l in t.splitlines() if not l.startswith("#")]
    assert L.povu_graph_vertex_count(b) == 2 * len(src["ids"]) and L.povu_flubbles_count(f2) == 2 * nfl1
    assert body(text2) == body(golden("minimal_substitution"))       # a flubble is called only where every reference path runs: component 1
    L.povu_flubbles_free(f2)
    f2, text2 = _vcf_via_ffi(L, b, ["ZZ1"])
    assert body(text2) == body(golden("minimal_substitution"))       # the new component, against its own reference
    L.povu_flubbles_free(f2); L.povu_graph_free(b)
    # vertices added in any order (sorted by id at finalize): same VCF as the id-ordered build
    b = L.povu_graph_new(0, 0, 0)
    order = list(range(len(src["ids"])))[::-1]
    for k in order:
        assert L.povu_graph_add_vertex(b, int(src["ids"][k]), src["seqs"][k].encode()) != C.c_size_t(-1).value
    for v1, s1, v2, s2 in zip(src["v1"], src["s1"], src["v2"], src["s2"]):
        assert L.povu_graph_add_edge(b, int(src["ids"][v1]), int(s1), int(src["ids"][v2]), int(s2)) != C.c_size_t(-1).value
    for nm, steps, strands in src["paths"]:
        arr = (PovuStep * len(steps))(*[PovuStep(int(src["ids"][v]), int(o)) for v, o in zip(steps, strands)])
        assert L.povu_graph_add_path(b, nm.encode(), arr, len(steps))
    f3, text3 = _vcf_via_ffi(L, b, ["HG1"])
    assert nodate(text3) == golden("minimal_substitution")
    L.povu_flubbles_free(f3); L.povu_graph_free(b)
    # one-shot file API
    out = os.path.join(tmp, "o.vcf"); refs = os.path.join(tmp, "r.txt"); open(refs, "w").write("HG1\n")
    err = PovuError()
    assert L.povu_gfa_to_vcf(helpers.fixture_path("nested_deletion").encode(), out.encode(), refs.encode(), C.byref(err)), err.message
    assert nodate(open(out).read()) == golden("nested_deletion")
    assert not L.povu_gfa_to_vcf(helpers.fixture_path("unsupported_overlap").encode(), out.encode(), refs.encode(), C.byref(err)) and err.code == 1
    L.povu_error_free(C.byref(err))
    # independent graphs on different threads (builder_tests.rs:106-137)
    res = {}
    def work(k, name, pref):
        gg = L.povu_graph_from_gfa(helpers.fixture_path(name).encode(), None)
        ff, tx = _vcf_via_ffi(L, gg, [pref])
        res[k] = nodate(tx) == golden(name)
        L.povu_flubbles_free(ff); L.povu_graph_free(gg)
    th = [threading.Thread(target=work, args=(k, n, p)) for k, (n, p) in enumerate([("chain_s3", "S001"), ("nested_s2_shuffled", "S001"), ("deletion_flubble", "HG1")])]
    [t.start() for t in th]; [t.join() for t in th]
    assert res == {0: True, 1: True, 2: True}


def check_ffi_perturbed(L, tmp, seeds):
    """povu_ffi builder path (povu_graph_new / add_vertex / add_edge / add_path -> find_flubbles -> call_variants -> vcf_to_string)
    on perturbed bubble graphs with random-walk haplotypes, whole VCF against the compiled reference's gfa2vcf"""
    import call_checks as cc
    from povu_b200 import synth
    cc.require_ref()
    n = 0
    for seed in seeds:
        g = cc.perturbed_graph(seed)
        gfa = os.path.join(tmp, f"ffi{seed}.gfa")
        synth.write_gfa(g, gfa)
        prefix = g["paths"][0][0].split("#")[0]
        r = subprocess.run([cc.REF, "gfa2vcf", "-i", gfa, "-t", "1", "-P", prefix], capture_output=True, text=True)
        if r.returncode != 0:
            continue
        b = L.povu_graph_new(0, 0, 0)
        for i, s in zip(g["ids"], g["seqs"]):
            assert L.povu_graph_add_vertex(b, int(i), s.encode()) != C.c_size_t(-1).value
        for v1, s1, v2, s2 in zip(g["v1"], g["s1"], g["v2"], g["s2"]):
            assert L.povu_graph_add_edge(b, int(g["ids"][v1]), int(s1), int(g["ids"][v2]), int(s2)) != C.c_size_t(-1).value
        for nm, steps, strands in g["paths"]:
            arr = (PovuStep * len(steps))(*[PovuStep(int(g["ids"][v]), int(o)) for v, o in zip(steps, strands)])
            assert L.povu_graph_add_path(b, nm.encode(), arr, len(steps))
        L.povu_graph_finalize(b)
        err = PovuError()
        arr = (C.c_char_p * 1)(prefix.encode())
        assert L.povu_graph_set_references_from_prefixes(b, arr, 1, C.byref(err))
        f = L.povu_graph_find_flubbles(b, C.byref(err))
        assert f, err.message
        o = L.povu_flubbles_call_variants(f, C.byref(err))
        if not o:      # the product refuses inputs on which the reference has undefined behaviour
            assert b"the reference aborts" in (err.message or b"") or b"reference" in (err.message or b""), err.message
            L.povu_flubbles_free(f); L.povu_graph_free(b)
            continue
        k = C.c_size_t()
        p = L.povu_vcf_to_string(o, C.byref(k))
        text = C.string_at(p, k.value).decode()
        L.povu_string_free(p); L.povu_vcf_free(o); L.povu_flubbles_free(f); L.povu_graph_free(b)
        a, m = nodate(r.stdout).splitlines(), nodate(text).splitlines()
        if a != m:      # only a clamped SUBR slice may differ (the reference reads past the end of a walk there: undefined)
            diff = [l for l in a if l not in set(m)] + [l for l in m if l not in set(a)]
            assert all("VARTYPE=SUBR" in l or "VARTYPE=" not in l for l in diff), (seed, diff[:2])
            a = [l for l in a if l.startswith("#") or ("VARTYPE=" in l and "VARTYPE=SUBR" not in l)]; m = [l for l in m if "VARTYPE=SUBR" not in l]
        assert a == m, seed
        n += 1
    return n
