"""Full-size runs on the B200 with a bit-exact sample check against the compiled reference.

For every configuration: generate the graph, run the whole hot path on the GPU, then cut a prefix of the
(chain-shaped) graph, run the UNMODIFIED reference (oracle/_ref/povu_ref) on that prefix and require that its
records are exactly the GPU records whose context lies inside the prefix (same text: ids, POS, alleles, genotypes).
usage: python scripts/scale_run.py cfg2|cfg3|cfg3s|cfg5s|cfg5h [...]"""
import json
import os
import re
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from povu_b200 import api, synth   # noqa: E402

REF = os.path.join(ROOT, "oracle", "_ref", "povu_ref")


def build(cfg):
    if cfg == "cfg2":
        g = synth.snp_indel_chain(400_000, n_paths=16, seed=42)
        return g, "S001#1", 20_000 * 5 // 2   # prefix: about 20 k bubbles (backbone vertex b_i has index base[i])
    if cfg in ("cfg3", "cfg3s"):
        inst, paths = (19_000, 90) if cfg == "cfg3" else (1_500, 24)
        T = synth.cfg3_templates(tower_depth=1024 if cfg == "cfg3" else 256)
        g = synth.tiled(T, inst, paths, seed=3)
        return g, "S001#1", None
    if cfg == "cfg5s":
        g = synth.snp_indel_chain(200_000, n_paths=232, seed=5)
        return g, "S001#1", 10_000 * 5 // 2
    if cfg == "cfg5h":   # half of cfg5's flubbles at its full path count: 500 k bubbles x 464 paths, ~0.43 G steps
        g = synth.snp_indel_chain(500_000, n_paths=464, seed=5)
        return g, "S001#1", 4_000 * 5 // 2
    raise SystemExit("unknown config " + cfg)


def prefix_cut(g, want):
    """a vertex index at which the chain can be cut: the first step index >= want of path 0 that every path visits"""
    import numpy as np
    common = None
    for _, st, _ in g["paths"]:
        s = np.unique(st[(st >= want) & (st < want + 5000)])
        common = s if common is None else np.intersect1d(common, s)
    return int(common[0])


def main():
    out = {}
    for cfg in sys.argv[1:]:
        t0 = time.time()
        g, prefix, cut_hint = build(cfg)
        V, E = len(g["ids"]), len(g["v1"])
        S = sum(len(p[1]) for p in g["paths"])
        t_gen = time.time() - t0
        pv = api.Povu(device=0) if not os.environ.get("PV_EMU_LIB") else api.Povu(lib_path=os.environ["PV_EMU_LIB"])
        t0 = time.time()
        s = api.Session.from_graph(pv, g)
        t_sess = time.time() - t0
        for _ in range(int(os.environ.get("SCALE_ITERS", "4"))):           # the page-locked result vectors and the device pool reach their steady state after 3 calls
            ms = s.hot_path([prefix])
        prof_top = None
        if os.environ.get("SCALE_PROF"):
            pv.set_profiling(True)
            s.hot_path([prefix])
            prof = pv.profile()
            pv.set_profiling(False)
            tot = sum(k["ms"] for k in prof)
            prof_top = [{"name": k["name"], "launches": k["launches"], "ms": round(k["ms"], 3), "share": round(k["ms"] / tot, 3),
                         "GBps": round(k["bytes"] / (k["ms"] * 1e-3) / 1e9, 1) if k["ms"] > 0 else None} for k in sorted(prof, key=lambda k: -k["ms"])[:14]]
        st = s.call_stats()
        nrec = int(s.lib.pvgpu_session_record_count(s.h)); nfl = int(s.lib.pvgpu_session_flubble_count(s.h))
        assert st["hard_errors"] == 0
        # sample check against the reference on a prefix
        cut = prefix_cut(g, cut_hint if cut_hint else V // 50)
        sub = synth.truncate(g, cut + 1)
        d = tempfile.mkdtemp()
        gfa = os.path.join(d, "p.gfa")
        synth.write_gfa_fast(sub, gfa)
        t0 = time.time()
        r = subprocess.run([REF, "gfa2vcf", "-i", gfa, "-P", prefix, "-t", "1"], capture_output=True, text=True)
        t_ref = time.time() - t0
        assert r.returncode == 0, r.stderr[-500:]
        lim = int(g["ids"][cut]) - 400          # stay clear of the cut itself
        def inside(line):
            f = line.split("\t")
            ids = [int(x) for x in re.findall(r"\d+", f[2])]
            return max(ids) < lim and "VARTYPE=SUBR" not in line and "TANGLED=T" not in line
        t0 = time.time()
        lines, _ = s.record_lines()
        t_text = time.time() - t0
        got = sorted(l.rstrip("\n") for l in lines if inside(l))
        want = sorted(l for l in r.stdout.splitlines() if not l.startswith("#") and inside(l))
        ok = got == want and len(want) > 0
        out[cfg] = {"V": V, "E": E, "paths": len(g["paths"]), "steps": S, "flubbles": nfl, "records": nrec, "ms_upload_dec_index_call": ms, "stats": st,
                    "sample_check": {"prefix_vertices": cut + 1, "reference_records_compared": len(want), "bit_exact": ok, "reference_s": t_ref},
                    "edges_per_s": E / ((ms[1] + ms[2] + ms[3]) * 1e-3), "records_per_s": nrec / ((ms[2] + ms[3]) * 1e-3),
                    "top_kernels": prof_top, "gen_s": t_gen, "session_s": t_sess, "text_s": t_text, "peak_device_GB": pv.peak_bytes / 1e9}
        print(cfg, json.dumps(out[cfg]), flush=True)
        assert ok, "sample differs from the reference"
        s.close(); pv.close()
    json.dump(out, open(os.path.join(ROOT, "gpurun_out", "scale_run.json"), "w"), indent=1)


if __name__ == "__main__":
    main()
