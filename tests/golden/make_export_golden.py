"""Regenerates tests/golden/export/ (run in the build container only): what the UNMODIFIED reference (oracle/_ref/povu_ref)
writes for `gfa2vcf --structure-export` (the JSON document and its .flubble-debug.jsonl side file) and prints for
`decompose -h` ("Boundary: b1 b2" lines on stderr) on the reference's own fixtures, two synthetic families and three small
arbitrary bidirected graphs whose DFS trees have bridges (hairpin boundaries) and capping back edges."""
import os
import shutil
import subprocess
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
BIN = os.path.join(ROOT, "oracle", "_ref", "povu_ref")
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(HERE))
import helpers  # noqa: E402
from povu_b200 import synth  # noqa: E402

OUT = os.path.join(HERE, "export")
WILD = {"wild_h36": 36, "wild_h46": 46, "wild_h59": 59}
CASES = [(n, "ref" if n == "hairpin_inversion_subr" else "HG1") for n in helpers.POSITIVE] + [("nested_s1", "S001"), ("chain_s3", "S001")] + [(n, "ref") for n in WILD]


def wild_graph(seed, path):
    """synth.wild(seed) plus one single-step path `ref#1#c` (gfa2vcf needs a reference to select)"""
    V = 10 + seed % 30
    g = synth.wild(seed, V=V, E=int(V * (1.1 + 0.1 * (seed % 7))), n_comp=1)
    synth.write_gfa(g, path)
    with open(path, "a") as f:
        f.write("P\tref#1#c\t%d+\t*\n" % g["ids"][0])


def main():
    os.makedirs(OUT, exist_ok=True)
    for name, seed in WILD.items():
        wild_graph(seed, os.path.join(HERE, "synthetic", name + ".gfa"))
    for name, prefix in CASES:
        gfa = helpers.fixture_path(name)
        d = tempfile.mkdtemp()
        js = os.path.join(d, "x.json")
        r = subprocess.run([BIN, "gfa2vcf", "-i", gfa, "-t", "1", "-P", prefix, "--structure-export", js], capture_output=True, text=True)
        assert r.returncode == 0, (name, r.stderr[-300:])
        import export_checks
        with open(os.path.join(OUT, name + ".structure.json"), "w") as f:
            f.write(export_checks.trim_steps(open(js).read(), gfa))      # (the reference over-reads the step arrays: see trim_steps)
        shutil.copy(js + ".flubble-debug.jsonl", os.path.join(OUT, name + ".frames.jsonl"))
        r = subprocess.run([BIN, "decompose", "-i", gfa, "-t", "1", "-h", "-o", d], capture_output=True, text=True)
        assert r.returncode == 0, (name, r.stderr[-300:])
        with open(os.path.join(OUT, name + ".hairpins.txt"), "w") as f:
            f.writelines(l + "\n" for l in r.stderr.splitlines() if l.startswith("Boundary:"))
        shutil.rmtree(d)


if __name__ == "__main__":
    main()
