"""Regenerates tests/golden/ from the reference (run in the build container only; the GPU box never
reads /root/reference).  Inputs: the reference's own fixture GFAs (test data, copied verbatim) and
tests/data/LPA.gfa (gzipped).  Expected outputs: produced by oracle/_ref/povu_ref, i.e. the UNMODIFIED
reference sources compiled by oracle/Makefile.  doc_rows.json / cycle_oracle.json / pvst_test.json are
hand-transcribed from the reference's own tests and docs (sources cited inside)."""
import gzip
import os
import shutil
import subprocess
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = "/root/reference"
BIN = os.path.join(ROOT, "oracle", "_ref", "povu_ref")
sys.path.insert(0, ROOT)

PREFIX = {"hairpin_inversion_subr": "ref", "LPA": "chm13"}


def run_ref(gfa, name):
    exp = os.path.join(HERE, "expected")
    os.makedirs(exp, exist_ok=True)
    d = tempfile.mkdtemp()
    r = subprocess.run([BIN, "decompose", "-i", gfa, "-o", d], stderr=subprocess.DEVNULL)
    if r.returncode == 0:
        for f in sorted(os.listdir(d)):
            shutil.copy(os.path.join(d, f), os.path.join(exp, f"{name}.{f}"))
    shutil.rmtree(d)
    r = subprocess.run([BIN, "gfa2vcf", "-i", gfa, "-t", "1", "-P", PREFIX.get(name, "HG1")], capture_output=True, text=True)
    with open(os.path.join(exp, f"{name}.rc"), "w") as f:
        f.write(f"{0 if r.returncode == 0 else 1}\n")
    if r.returncode == 0:
        lines = [l for l in r.stdout.splitlines(True) if not l.startswith("##fileDate")]
        with open(os.path.join(exp, f"{name}.vcf"), "w") as f:
            f.writelines(lines)
    else:
        with open(os.path.join(exp, f"{name}.stderr"), "w") as f:
            f.write("Invalid GFA\n" if "Invalid GFA" in r.stderr else r.stderr[-200:])


def main():
    fx = os.path.join(HERE, "fixtures")
    os.makedirs(fx, exist_ok=True)
    src = os.path.join(REF, "tests", "lean4_conformance", "fixtures")
    for f in sorted(os.listdir(src)):
        if f.endswith(".gfa"):
            shutil.copy(os.path.join(src, f), os.path.join(fx, f))
            run_ref(os.path.join(fx, f), f[:-4])
    lpa = os.path.join(REF, "tests", "data", "LPA.gfa")
    with open(lpa, "rb") as i, gzip.GzipFile(os.path.join(fx, "LPA.gfa.gz"), "wb", mtime=0) as o:
        shutil.copyfileobj(i, o)
    run_ref(lpa, "LPA")
    # synthetic families pinned against the compiled reference (small, deterministic)
    from povu_b200 import synth
    syn = os.path.join(HERE, "synthetic")
    os.makedirs(syn, exist_ok=True)
    cases = {"nested_s1": synth.nested_bubbles(1, n_top=12, max_depth=3, n_paths=6),
             "nested_s2_shuffled": synth.nested_bubbles(2, n_top=12, max_depth=3, n_paths=6, shuffle_links=True),
             "chain_s3": synth.snp_indel_chain(60, n_paths=6, seed=3),
             "multi_comp": synth.concat_components([synth.nested_bubbles(4, n_top=5), synth.snp_indel_chain(10, 4, seed=5), synth.nested_bubbles(6, n_top=4)])}
    for name, g in cases.items():
        p = os.path.join(syn, name + ".gfa")
        synth.write_gfa(g, p)
        PREFIX[name] = "S001#1#chr1" if name == "multi_comp" else "S001"
        run_ref(p, name)


if __name__ == "__main__":
    main()
