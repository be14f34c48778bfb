"""Shared test helpers: a tiny GFA reader (numpy arrays, same layout as povu_b200.synth) and golden access."""
import gzip
import json
import os

import numpy as np

from povu_b200 import synth

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
POSITIVE = ["minimal_substitution", "insertion_flubble", "deletion_flubble", "nested_deletion",
            "nested_substitution_missing_outer", "repeat_anchor_deletion", "repeat_anchor_insertion",
            "complex_substitution_span", "hairpin_inversion_subr", "linear_no_variant", "two_ordered_substitutions"]
NEGATIVE = ["unsupported_overlap", "malformed_path_missing_overlaps"]
SYNTHETIC = ["nested_s1", "nested_s2_shuffled", "chain_s3", "multi_comp"]


def fixture_path(name, tmpdir=None):
    if name == "LPA":
        gz = os.path.join(GOLDEN, "fixtures", "LPA.gfa.gz")
        out = os.path.join(tmpdir or "/tmp", "LPA.gfa")
        if not os.path.exists(out):
            with gzip.open(gz, "rb") as i, open(out, "wb") as o:
                o.write(i.read())
        return out
    for sub in ("fixtures", "synthetic"):
        p = os.path.join(GOLDEN, sub, name + ".gfa")
        if os.path.exists(p):
            return p
    raise FileNotFoundError(name)


def read_gfa(path):
    segs, L, P = {}, [], []
    with open(path) as fh:
        for line in fh:
            f = line.rstrip("\n").split("\t")
            if f[0] == "S":
                segs.setdefault(int(f[1]), f[2])
            elif f[0] == "L":
                L.append((int(f[1]), f[2], int(f[3]), f[4]))
            elif f[0] == "P":
                P.append((f[1], f[2]))
    ids = np.array(sorted(segs), dtype=np.uint32)
    idx = {int(x): i for i, x in enumerate(ids)}
    links = [synth.link(idx[a], oa, idx[b], ob) for a, oa, b, ob in L]
    paths = []
    for name, steps in P:
        st = steps.split(",")
        paths.append((name, np.array([idx[int(s[:-1])] for s in st], dtype=np.uint32), np.array([0 if s[-1] == "+" else 1 for s in st], dtype=np.uint8)))
    return synth.make_graph(ids, [segs[int(i)] for i in ids], links, paths)


def expected_pvsts(name):
    exp = os.path.join(GOLDEN, "expected")
    out = {}
    for f in os.listdir(exp):
        if f.startswith(name + ".") and f.endswith(".pvst"):
            out[int(f[len(name) + 1:-5])] = open(os.path.join(exp, f)).read()
    return out


def load_json(name):
    with open(os.path.join(GOLDEN, name)) as f:
        return json.load(f)
