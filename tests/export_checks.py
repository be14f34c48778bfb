"""`--hairpins` and `--structure-export` of the CLI clone against (a) the committed outputs of the compiled
reference (tests/golden/export/, made by tests/golden/make_export_golden.py) and (b) the compiled reference itself on seeded
graphs.  Shared by the host-rehearsal tests (CPU) and the GPU tests; `cli` is the povu binary under test.

What is compared byte for byte: the "Boundary: b1 b2" stderr lines of `decompose -h` (flubbles.cpp:711-717), the
<json>.flubble-debug.jsonl side file (tree-edge indices / ids and new_class() ids of every class-stack entry,
flubbles.cpp:108-213) and, for single-tree graphs, the whole JSON document (to_structure_export.cpp).  With several PVST trees
the reference numbers them in directory-iteration order of its temporary forest directory (unspecified), so those documents are
compared after renaming the trees canonically."""
import json
import os
import subprocess

import helpers
from povu_b200 import synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.path.join(ROOT, "oracle", "_ref", "povu_ref")
EXPORT = os.path.join(helpers.GOLDEN, "export")
GOLDEN_CASES = [(n, "ref" if n == "hairpin_inversion_subr" else "HG1") for n in helpers.POSITIVE] + [("nested_s1", "S001"), ("chain_s3", "S001")] + \
               [(n, "ref") for n in ("wild_h36", "wild_h46", "wild_h59")]


def run(*cmd):
    return subprocess.run(list(cmd), capture_output=True, text=True)


def boundaries(stderr):
    return [l for l in stderr.splitlines() if l.startswith("Boundary:") or "WARN: curr boundary" in l]


def read(p):
    with open(p) as f:
        return f.read()


def check_call_export(cli, tmp):
    """`povu call --structure-export` (no decompose in the run): the frames are whatever the side file already holds
    (to_structure_export.cpp:146-170) -- none without one, its lines verbatim with one"""
    assert os.path.exists(REF)
    gfa = helpers.fixture_path("nested_deletion")
    d = os.path.join(tmp, "callexp"); os.makedirs(d)
    assert run(REF, "decompose", "-i", gfa, "-o", d, "-t", "1").returncode == 0
    rj, mj = os.path.join(tmp, "callexp.r.json"), os.path.join(tmp, "callexp.m.json")
    for side in (None, '{"schema":"anything"}\n'):
        for j in (rj, mj):
            if side is None:
                if os.path.exists(j + ".flubble-debug.jsonl"):
                    os.remove(j + ".flubble-debug.jsonl")
            else:
                with open(j + ".flubble-debug.jsonl", "w") as f:
                    f.write(side)
        r = run(REF, "call", "-i", gfa, "-f", d, "-t", "1", "-P", "HG1", "--structure-export", rj)
        m = run(cli, "call", "-i", gfa, "-f", d, "-t", "1", "-P", "HG1", "--structure-export", mj)
        assert r.returncode == 0 and m.returncode == 0, m.stderr[-300:]
        assert trim_steps(read(rj), gfa) == read(mj)
        assert len(json.loads(read(mj))["flubble_debug"]["frames"]) == (0 if side is None else 1)


def trim_steps(doc, gfa):
    """The reference lists `hap_len` (base pairs) entries of a path's step arrays instead of `step_count`
    (to_structure_export.cpp:319-328: `step_idx < ref.get_length()`), i.e. it reads past the end of the walk whenever a
    segment is longer than one base; what lies there is whatever the allocator left.  The defined part is the first
    step_count entries: cut the reference's arrays there (text surgery, everything else stays byte for byte)."""
    counts = []
    with open(gfa) as f:
        for line in f:
            if line.startswith("P\t"):
                counts.append(len(line.split("\t")[2].split(",")))
    out, pos, k = [], 0, 0
    key = '"steps":['
    while True:
        i = doc.find(key, pos)
        if i < 0:
            break
        j = doc.index("]", i)
        items = doc[i + len(key):j].split(",") if j > i + len(key) else []
        out.append(doc[pos:i + len(key)] + ",".join(items[:counts[k]]))
        pos = j; k += 1
    out.append(doc[pos:])
    return "".join(out)


def check_export_golden(cli, tmp):
    for name, prefix in GOLDEN_CASES:
        gfa = helpers.fixture_path(name)
        js = os.path.join(tmp, name + ".json")
        r = run(cli, "gfa2vcf", "-i", gfa, "-t", "1", "-P", prefix, "--structure-export", js)
        assert r.returncode == 0, (name, r.stderr[-400:])
        assert read(js + ".flubble-debug.jsonl") == read(os.path.join(EXPORT, name + ".frames.jsonl")), name
        assert read(js) == read(os.path.join(EXPORT, name + ".structure.json")), name
        json.loads(read(js))
        d = os.path.join(tmp, name + ".forest"); os.makedirs(d, exist_ok=True)
        r = run(cli, "decompose", "-i", gfa, "-h", "-o", d)
        assert r.returncode == 0, (name, r.stderr[-400:])
        assert "".join(l + "\n" for l in boundaries(r.stderr)) == read(os.path.join(EXPORT, name + ".hairpins.txt")), name
    # the reference's own integration test for this surface (tests/integration_tests/gfa2vcf_tests.cc:137-166,
    # StructureExportIncludesFlubbleDebugFrame): exit 0, the four keys in the document, the side file exists
    gfa = helpers.fixture_path("insertion_flubble"); js = os.path.join(tmp, "ref_test.json")
    r = run(cli, "gfa2vcf", "-i", gfa, "-t", "1", "-P", "HG1", "--structure-export", js)
    assert r.returncode == 0, r.stderr
    text = read(js)
    for key in ('"flubble_debug"', '"stack_entries"', '"next_seen_table"', '"diagnostic":"ok"'):
        assert key in text, key
    assert os.path.exists(js + ".flubble-debug.jsonl")
    # the VCF on stdout is not disturbed by the export, and a second run replaces (not appends to) the side file
    gfa = helpers.fixture_path("nested_deletion"); js = os.path.join(tmp, "twice.json")
    a = run(cli, "gfa2vcf", "-i", gfa, "-P", "HG1", "--structure-export", js)
    b = run(cli, "gfa2vcf", "-i", gfa, "-P", "HG1", "--structure-export=" + js)
    c = run(cli, "gfa2vcf", "-i", gfa, "-P", "HG1")
    assert a.stdout == b.stdout == c.stdout and read(js + ".flubble-debug.jsonl") == read(os.path.join(EXPORT, "nested_deletion.frames.jsonl"))


def check_cycle_tables(cli, tmp):
    """the reference's conformance harness reads the debug frames of `--structure-export` and checks six hand-written cycle tables
    (tests/lean4_conformance/src/main.rs:1594-1622, transcribed in tests/golden/cycle_oracle.json): tree-edge ids in one group
    share a class, ids of different groups do not.  Here the same check on the product's frames."""
    tables = helpers.load_json("cycle_oracle.json")
    for name, groups in tables.items():
        if name.startswith("_"):
            continue
        js = os.path.join(tmp, "cyc_" + name + ".json")
        r = run(cli, "gfa2vcf", "-i", helpers.fixture_path(name), "-t", "1", "-P", "HG1", "--structure-export", js)
        assert r.returncode == 0, (name, r.stderr[-300:])
        frame = json.loads(read(js))["flubble_debug"]["frames"][0]
        cls = {e["tree_edge_id"]: e["class_id"] for e in frame["stack_entries"]}
        assert all(e["diagnostic"] == "ok" for e in frame["stack_entries"]), name
        seen = {}
        for gi, grp in enumerate(groups):
            ids = {cls[t] for t in grp}
            assert len(ids) == 1, (name, grp, "one class expected")
            c = ids.pop()
            assert c not in seen, (name, grp, "distinct groups must not share a class")
            seen[c] = gi


def check_hierarchy_spans(cli, tmp):
    """the conformance harness's hierarchy rule (tests/lean4_conformance/src/main.rs:1490-1571), restated: every class-stack
    span (an entry and the next later entry of its class, more than one position apart) names a PVST flubble node by its
    canonical boundary id, and that node's parent is the node of the smallest strictly containing span, else the tree's root."""
    cases = [(n, "ref" if n == "hairpin_inversion_subr" else "HG1") for n in helpers.POSITIVE] + [("nested_s1", "S001"), ("nested_s2_shuffled", "S001"), ("chain_s3", "S001")]
    checked = 0
    for name, prefix in cases:
        js = os.path.join(tmp, "hier_" + name + ".json")
        r = run(cli, "gfa2vcf", "-i", helpers.fixture_path(name), "-t", "1", "-P", prefix, "--structure-export", js)
        assert r.returncode == 0, (name, r.stderr[-300:])
        d = json.loads(read(js))
        by_boundary = {n["boundary"]: n for n in d["pvst_nodes"] if n["family"] == "flubble"}
        assert len(by_boundary) == sum(1 for n in d["pvst_nodes"] if n["family"] == "flubble"), (name, "duplicate boundary id")
        for frame in d["flubble_debug"]["frames"]:
            ent = frame["stack_entries"]
            spans = []
            for i, e in enumerate(ent):
                close = next((j for j in range(i + 1, len(ent)) if ent[j]["class_id"] == e["class_id"]), None)
                if close is None or i + 1 >= close:
                    continue
                c = ent[close]
                bid = f">{c['boundary_vertex_id']}>{e['boundary_vertex_id']}" if e["orientation"] == "<" and c["orientation"] == "<" else \
                    f"{e['orientation']}{e['boundary_vertex_id']}{c['orientation']}{c['boundary_vertex_id']}"
                assert bid in by_boundary, (name, bid, "a stack span without a PVST flubble node")
                spans.append((e["order"], c["order"], by_boundary[bid]))
            for a, b, node in spans:
                cont = [(b2 - a2, n2["node_id"]) for a2, b2, n2 in spans if n2["node_id"] != node["node_id"] and a2 <= a and b <= b2 and (a2 < a or b < b2)]
                want = min(cont)[1] if cont else node["node_id"].split(":")[0] + ":0"
                assert node["parent"] == want, (name, node["node_id"], node["parent"], want)
                checked += 1
    assert checked > 100, checked


def canon(doc):
    """tree numbers -> names that do not depend on the order the trees were read in"""
    d = json.loads(doc)
    name = {}
    for n in d["pvst_nodes"]:
        name.setdefault(n["component"], []).append(n["label"])
    name = {k: "|".join(v) for k, v in name.items()}
    ren = lambda nid: None if nid is None else name[int(nid.split(":")[0])] + ":" + nid.split(":")[1]
    for n in d["pvst_nodes"]:
        n["node_id"] = ren(n["node_id"]); n["parent"] = ren(n["parent"]); n["children"] = [ren(c) for c in n["children"]]
        n.pop("order"); n.pop("component")
    for n in d["boundary_candidates"]:
        n["node_id"] = ren(n["node_id"]); n.pop("order")
    for n in d["variant_calls"]:
        n["source"]["node_id"] = ren(n["source"]["node_id"])
    d["pvst_nodes"].sort(key=lambda n: n["node_id"]); d["boundary_candidates"].sort(key=lambda n: n["node_id"])
    d["flubble_debug"]["frames"].sort(key=lambda f: json.dumps(f, sort_keys=True))
    return d


def compare_with_reference(cli, gfa, prefix, tmp, tag, chunk=None):
    """returns (#stack entries compared, #variant calls compared); the reference may crash on arbitrary graphs: then only the
    frames it managed to write are compared"""
    assert os.path.exists(REF), "oracle/_ref/povu_ref is missing (python __graft_entry__.py builds it where /root/reference exists)"
    rj, mj = os.path.join(tmp, tag + ".ref.json"), os.path.join(tmp, tag + ".mine.json")
    extra = ["-c", str(chunk)] if chunk else []
    r = run(REF, "gfa2vcf", "-i", gfa, "-t", "1", "-P", prefix, "--structure-export", rj, *extra)
    m = run(cli, "gfa2vcf", "-i", gfa, "-t", "1", "-P", prefix, "--structure-export", mj, *extra)
    if not os.path.exists(rj + ".flubble-debug.jsonl"):      # no component with three or more vertices (or the reference crashed early)
        if r.returncode == 0:
            assert m.returncode == 0 and not os.path.exists(mj + ".flubble-debug.jsonl") and trim_steps(read(rj), gfa) == read(mj), (tag, "document without frames")
        return 0, 0
    rf, mf = read(rj + ".flubble-debug.jsonl"), read(mj + ".flubble-debug.jsonl") if os.path.exists(mj + ".flubble-debug.jsonl") else ""
    if r.returncode != 0:
        assert mf.startswith(rf), (tag, "frames (reference crashed later)")
        return rf.count('"order"'), 0
    assert m.returncode == 0, (tag, m.stderr[-400:])
    assert rf == mf, (tag, "frames")
    a, b = trim_steps(read(rj), gfa), read(mj)
    if len(json.loads(a)["flubble_debug"]["frames"]) <= 1:
        assert a == b, (tag, "document")
    else:
        assert canon(a) == canon(b), (tag, "document (canonical tree names)")
    return rf.count('"order"'), len(json.loads(a)["variant_calls"])


def compare_hairpins(cli, gfa, tmp, tag):
    d = os.path.join(tmp, tag + ".d"); os.makedirs(d, exist_ok=True)
    r = run(REF, "decompose", "-i", gfa, "-t", "1", "-h", "-o", d)
    if r.returncode != 0:
        return None
    m = run(cli, "decompose", "-i", gfa, "-h", "-o", d)
    assert m.returncode == 0, (tag, m.stderr[-400:])
    assert boundaries(m.stderr) == boundaries(r.stderr), tag
    return len(boundaries(r.stderr))


def wild_with_ref(seed, path, V=None, n_comp=1):
    V = V or 10 + seed % 30
    g = synth.wild(seed, V=V, E=int(V * (1.1 + 0.1 * (seed % 7))), n_comp=n_comp)
    synth.write_gfa(g, path)
    with open(path, "a") as f:
        f.write("P\tref#1#c\t%d+\t*\n" % g["ids"][0])


def check_wild_vs_reference(cli, tmp, seeds, V=None):
    """arbitrary bidirected multigraphs (self-loops, parallel links, bridges, several components): hairpin report + frames"""
    nb = ne = 0
    for seed in seeds:
        gfa = os.path.join(tmp, "wild%d.gfa" % seed)
        wild_with_ref(seed, gfa, V=V, n_comp=1 + seed % 3)
        k = compare_hairpins(cli, gfa, tmp, "wild%d" % seed)
        nb += k or 0
        ne += compare_with_reference(cli, gfa, "ref", tmp, "wild%d" % seed)[0]
    return nb, ne


DEGENERATE = {
    "one": "S\t1\tA\nP\tref#1#c\t1+\t*\n",
    "two": "S\t1\tA\nS\t2\tC\nL\t1\t+\t2\t+\t0M\nP\tref#1#c\t1+,2+\t*\n",
    "three": "S\t1\tA\nS\t2\tC\nS\t3\tG\nL\t1\t+\t2\t+\t0M\nL\t2\t+\t3\t+\t0M\nP\tref#1#c\t1+,2+,3+\t*\n",
    "selfloop": "S\t1\tA\nS\t2\tC\nS\t3\tG\nL\t1\t+\t2\t+\t0M\nL\t2\t+\t2\t+\t0M\nL\t2\t+\t3\t+\t0M\nP\tref#1#c\t1+,2+,2+,3+\t*\n",
    "ring": "S\t1\tA\nS\t2\tC\nS\t3\tG\nL\t1\t+\t2\t+\t0M\nL\t2\t+\t3\t+\t0M\nL\t3\t+\t1\t+\t0M\nP\tref#1#c\t1+,2+,3+\t*\n",
    "hairpin": "S\t1\tA\nS\t2\tC\nS\t3\tG\nS\t4\tT\nL\t1\t+\t2\t+\t0M\nL\t2\t+\t3\t+\t0M\nL\t2\t+\t3\t-\t0M\nL\t1\t-\t4\t+\t0M\nP\tref#1#c\t1+,2+,3+\t*\n",
    "mixed": "S\t1\tA\nS\t2\tC\nS\t3\tG\nS\t4\tT\nS\t5\tT\nS\t6\tAC\nS\t7\tT\nL\t1\t+\t2\t+\t0M\nL\t1\t+\t3\t+\t0M\nL\t2\t+\t4\t+\t0M\nL\t3\t+\t4\t+\t0M\nL\t5\t+\t6\t+\t0M\n"
             "P\tref#1#c\t1+,2+,4+\t*\nP\talt#1#c\t1+,3+,4+\t*\n",
}


def check_degenerate_vs_reference(cli, tmp):
    """a lone segment, a two-segment component (too small for the cycle-class pass: no frame, no side file), a bare chain, a
    self-loop, a ring without tips, a hairpin loop next to a tip, components of mixed sizes with an isolated segment"""
    for name, body in DEGENERATE.items():
        gfa = os.path.join(tmp, "deg_" + name + ".gfa")
        with open(gfa, "w") as f:
            f.write("H\tVN:Z:1.0\n" + body)
        assert compare_hairpins(cli, gfa, tmp, "deg_" + name) is not None, name
        compare_with_reference(cli, gfa, "ref", tmp, "deg_" + name)


def check_families_vs_reference(cli, tmp):
    """graphs with real variant calls: nested bubbles (diploid samples, missing genotypes), a VNTR-like repeat (TANGLED=T),
    inversion units (SUBR records), a deep tower, several components, small chunk sizes"""
    cases = {
        "nested": (synth.nested_bubbles(11, n_top=20, max_depth=4, n_paths=8), "S001", None),
        "nested_chunk3": (synth.nested_bubbles(12, n_top=15, max_depth=3, n_paths=6), "S001", 3),
        "repeat": (synth.repeat_loops(3, n_units=5, n_paths=6), "S001", None),
        "cyclic": (synth.cyclic_units(2, n_units=4, n_paths=6), "S001", None),
        "chain": (synth.snp_indel_chain(150, n_paths=6, seed=9), "S001", 7),
        "multi": (synth.concat_components([synth.nested_bubbles(4, n_top=5), synth.snp_indel_chain(10, 4, seed=5), synth.nested_bubbles(6, n_top=4)]), "S001#1#chr1", None),
    }
    tot = 0
    for tag, (g, prefix, chunk) in cases.items():
        gfa = os.path.join(tmp, tag + ".gfa")
        synth.write_gfa(g, gfa)
        ne, nv = compare_with_reference(cli, gfa, prefix, tmp, tag, chunk)
        assert ne > 0, tag
        tot += nv
    assert tot > 50
    # the reference's own inversion fixture (one SUBR record: source kind "hairpin") and LPA (2 238 records, 57 % tangled)
    for name, prefix in (("hairpin_inversion_subr", "ref"), ("LPA", "chm13")):
        ne, nv = compare_with_reference(cli, helpers.fixture_path(name, tmp), prefix, tmp, name)
        assert ne > 0 and nv > 0, name
