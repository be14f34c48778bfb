"""The CPU oracle (oracle/*.hpp) pinned against the reference's own golden vectors and against outputs
of the compiled reference (tests/golden/expected, made by tests/golden/make_golden.py)."""
import numpy as np
import pytest

import helpers
import oracle_binding as ob


@pytest.mark.parametrize("name", helpers.POSITIVE + helpers.SYNTHETIC + ["LPA"])
def test_oracle_pvst_matches_compiled_reference(built, name, tmp_path):
    rows = ob.oracle_decompose_gfa(helpers.fixture_path(name, str(tmp_path)))
    got = ob.rows_to_pvst_texts(rows)
    want = helpers.expected_pvsts(name)
    # components without flubbles still get a file (dummy root only)
    for cid, text in want.items():
        assert got.get(cid, "H\t0.0.3\t.\t.\t.\nD\t0\t.\t.\t.\n") == text


@pytest.mark.parametrize("name", ["minimal_substitution", "insertion_flubble", "deletion_flubble", "nested_deletion",
                                  "nested_substitution_missing_outer", "two_ordered_substitutions"])
def test_oracle_cycle_classes_match_reference_table(built, name):
    """tests/lean4_conformance/src/main.rs:1594-1622"""
    groups = helpers.load_json("cycle_oracle.json")[name]
    e, c, _, _ = ob.oracle_class_stack(helpers.fixture_path(name))
    cls = {int(a): int(b) for a, b in zip(e, c)}
    seen = {}
    for gi, grp in enumerate(groups):
        vals = {cls[x] for x in grp}
        assert len(vals) == 1, f"group {grp} split"
        v = vals.pop()
        assert v not in seen, f"groups {seen[v]} and {grp} merged"
        seen[v] = grp
    assert sorted(cls) == sorted(x for g in groups for x in g)


def test_oracle_pvst_known_answer(built):
    """tests/integration_tests/pvst_tests.cc:69-128"""
    t = helpers.load_json("pvst_test.json")
    ids = np.array(t["ids"], dtype=np.uint32)
    idx = {int(x): i for i, x in enumerate(ids)}
    g = {"ids": ids,
         "v1": np.array([idx[e[0]] for e in t["edges"]], dtype=np.uint32), "s1": np.array([0 if e[1] == "l" else 1 for e in t["edges"]], dtype=np.uint8),
         "v2": np.array([idx[e[2]] for e in t["edges"]], dtype=np.uint32), "s2": np.array([0 if e[3] == "l" else 1 for e in t["edges"]], dtype=np.uint8)}
    for infer in (False, True):
        rows = ob.oracle_decompose(g, infer_tips=infer)
        labels = [f"{'>' if af else '<'}{a}{'>' if zf else '<'}{z}" for _, a, af, z, zf, _ in rows]
        parents = ["." if p == 0 else labels[p - 1] for *_, p in rows]
        assert sorted(zip(labels, parents)) == sorted(map(tuple, t["expected"]))


def _vcf_rows(text):
    """VCF text -> rows shaped like tests/golden/doc_rows.json (the columns the reference's documents pin)."""
    samples, rows = [], []
    for l in text.splitlines():
        f = l.split("\t")
        if l.startswith("#CHROM"):
            samples = f[9:]
        elif not l.startswith("#"):
            rows.append({"CHROM": f[0], "POS": int(f[1]), "ID": f[2], "REF": f[3], "ALT": f[4], "INFO": f[7], "GT": f[9:]})
    return samples, rows


DOC = {k: v for k, v in helpers.load_json("doc_rows.json").items() if not k.startswith("_")}


@pytest.mark.parametrize("name", sorted(DOC))
def test_documented_golden_rows(built, name):
    """The golden VCF rows the reference documents (docs/vcf-modernization/practical_vcf_conformance.md:82-300, the same
    strings as povu-rs/tests/native_gfa_vcf_tests.rs:141-258) pin BOTH checkers: the committed outputs of the compiled
    reference (tests/golden/expected, what the CUDA path is compared with byte for byte) and the CPU oracle restatement."""
    import call_checks
    want = DOC[name]
    exp = open(helpers.GOLDEN + "/expected/" + name + ".vcf").read()
    samples, rows = _vcf_rows(exp)
    assert samples == want["samples"] and rows == want["rows"], "compiled reference output differs from the documented rows"
    if all("VARTYPE=SUBR" not in r["INFO"] for r in want["rows"]):      # the restatement does not cover the SNE pass
        samples, rows = _vcf_rows(call_checks.oracle_vcf(helpers.fixture_path(name), [want["prefix"]]))
        assert samples == want["samples"] and rows == want["rows"], "oracle restatement differs from the documented rows"
