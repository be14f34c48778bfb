#!/usr/bin/env python
"""`povu gfa2vcf` end to end on the cfg2 GFA text: our CLI (process start -> VCF file on disk) next to the UNMODIFIED
reference binary (oracle/_ref/povu_ref) on the same file, outputs compared byte for byte (##fileDate aside).
usage: cli_e2e.py [n_bubbles] [--no-ref]"""
import json
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from povu_b200 import synth  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 and sys.argv[1].isdigit() else 400_000
CLI = os.path.join(ROOT, "povu_b200", "bin", "povu")
REF = os.path.join(ROOT, "oracle", "_ref", "povu_ref")
PREFIX = "S001#1"
if "--lpa" in sys.argv:      # the reference's own real-data fixture (tests/data/LPA.gfa: 57 % of its regions are tangled)
    import gzip
    fp = "/tmp/cli_e2e_LPA.gfa"
    with gzip.open(os.path.join(ROOT, "tests", "golden", "fixtures", "LPA.gfa.gz"), "rb") as i, open(fp, "wb") as o:
        o.write(i.read())
    PREFIX = "chm13"
    out = {"workload": "tests/data/LPA.gfa (real data, tangled regions), -P chm13", "gfa_bytes": os.path.getsize(fp), "host_cores": os.cpu_count()}
else:
    fp = "/tmp/cli_e2e.gfa"
    g = synth.snp_indel_chain(n, n_paths=16, seed=42, with_seqs=True)
    synth.write_gfa_fast(g, fp)
    out = {"workload": f"snp_indel_chain({n} bubbles), 16 haplotype paths", "gfa_bytes": os.path.getsize(fp), "host_cores": os.cpu_count()}
open(fp, "rb").read()
runs = []
for i in range(4):
    t0 = time.perf_counter()
    with open("/tmp/cli_e2e_ours.vcf", "wb") as f:
        r = subprocess.run([CLI, "gfa2vcf", "-i", fp, "-P", PREFIX], stdout=f, stderr=subprocess.PIPE, env=dict(os.environ, POVU_B200_TIMING="1"))
    runs.append({"wall_s": time.perf_counter() - t0, "phases": r.stderr.decode().strip().splitlines()[-2:] if r.stderr else ""})
    assert r.returncode == 0, r.stderr.decode()
out["povu_b200"] = {"wall_s": min(x["wall_s"] for x in runs), "runs": runs, "vcf_bytes": os.path.getsize("/tmp/cli_e2e_ours.vcf")}
if "--no-ref" not in sys.argv and os.path.exists(REF):
    t0 = time.perf_counter()
    with open("/tmp/cli_e2e_ref.vcf", "wb") as f:
        r = subprocess.run([REF, "gfa2vcf", "-i", fp, "-P", PREFIX, "-t", "1"], stdout=f, stderr=subprocess.PIPE)
    out["reference"] = {"wall_s": time.perf_counter() - t0, "rc": r.returncode}
    strip = lambda p: [l for l in open(p, "rb").read().split(b"\n") if not l.startswith(b"##fileDate")]
    a, b = strip("/tmp/cli_e2e_ours.vcf"), strip("/tmp/cli_e2e_ref.vcf")
    out["identical_to_reference"] = a == b
    out["records"] = sum(1 for l in a if l and not l.startswith(b"#"))
    out["speedup_wall"] = out["reference"]["wall_s"] / out["povu_b200"]["wall_s"]
# `povu decompose`: .pvst forest on disk (device decompose + device .pvst text)
import shutil
import tempfile
d_ours, d_ref = tempfile.mkdtemp(), tempfile.mkdtemp()
runs = []
for i in range(3):
    t0 = time.perf_counter()
    r = subprocess.run([CLI, "decompose", "-i", fp, "-o", d_ours], capture_output=True)
    runs.append(time.perf_counter() - t0)
    assert r.returncode == 0, r.stderr.decode()
out["decompose_cli"] = {"povu_b200_wall_s": min(runs), "runs": runs, "pvst_bytes": sum(os.path.getsize(os.path.join(d_ours, f)) for f in os.listdir(d_ours))}
if "--no-ref" not in sys.argv and os.path.exists(REF):
    t0 = time.perf_counter()
    r = subprocess.run([REF, "decompose", "-i", fp, "-o", d_ref, "-t", "1"], capture_output=True)
    out["decompose_cli"]["reference_wall_s"] = time.perf_counter() - t0
    same = sorted(os.listdir(d_ours)) == sorted(os.listdir(d_ref)) and all(open(os.path.join(d_ours, f), "rb").read() == open(os.path.join(d_ref, f), "rb").read() for f in os.listdir(d_ref))
    out["decompose_cli"]["identical_to_reference"] = same
shutil.rmtree(d_ours); shutil.rmtree(d_ref)
print(json.dumps(out, indent=1))
