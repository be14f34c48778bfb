"""Full-size runs on the B200 with bit-exact window checks against the compiled reference.

For every configuration: generate the graph, run the whole hot path on the GPU, write ALL records as VCF text on the
device, then cut N_WINDOWS windows spread over the (chain-shaped) graph, run the UNMODIFIED reference
(oracle/_ref/povu_ref) on every window and require that its records are exactly the GPU records whose ids lie inside
the window -- every record type, SUBR and TANGLED=T lines included; same ID, REF, ALT, INFO, genotypes; POS shifted by
the bases of the reference walk before the window.
usage: python scripts/scale_run.py cfg2|cfg3|cfg3m|cfg5|cfg5h [...]      (results: gpurun_out/scale_run.json)"""
import json
import os
import re
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from povu_b200 import api, synth   # noqa: E402

REF = os.path.join(ROOT, "oracle", "_ref", "povu_ref")
N_WINDOWS = int(os.environ.get("SCALE_WINDOWS", "8"))


def cfg3_graph(instances, paths, tower_depth, seed=3):
    """nested bubble templates (depth mostly <= 3, some 4-8; 1 % inversion bubbles inside them), one
    `tower_depth`-deep PVST tower, VNTR-like repeat units (tangled regions) and blocks that 1 walk in 5 crosses on the reverse
    strand (whole-region inversions -> SNE pin pairs -> SUBR records); template frequencies below."""
    T = synth.cfg3_templates(seed=7, n_nested=24, tower_depth=tower_depth, p_inv=0.01, n_repeat=2)
    n_plain = len(T) - 1
    T += [synth.inversion_unit(k, inner=1) for k in range(4)]
    # nested + repeat templates: weight 1 each; the tower ~1 % of the instances; inversion blocks 1 % in total
    w = [1.0] * n_plain + [0.01 * n_plain] + [0.0025 * n_plain] * 4
    g = synth.tiled(T, instances, paths, seed=seed, tpl_weights=w)
    return g


def build(cfg):
    """-> graph, reference prefix, window width in instances"""
    if cfg == "cfg2":
        return synth.snp_indel_chain(400_000, n_paths=16, seed=42), "S001#1", 20_000
    if cfg == "cfg3":      # V ~ 10 M, 90 paths, 1024-deep tower
        return cfg3_graph(int(os.environ.get("CFG3_INSTANCES", "29000")), 90, 1024), "S001#1", 600
    if cfg == "cfg3m":     # the same mix at ~1 M segments (the size of the -m gpu test)
        return cfg3_graph(3_300, 90, 256), "S001#1", 150
    if cfg == "cfg5":      # call-heavy: 10^6 bubbles x 464 paths, ~0.87 G path steps
        return synth.snp_indel_chain(1_000_000, n_paths=464, seed=5), "S001#1", 2_000
    if cfg == "cfg5h":
        return synth.snp_indel_chain(500_000, n_paths=464, seed=5), "S001#1", 2_000
    raise SystemExit("unknown config " + cfg)


def check_windows(g, lines, prefix, width, n_windows=N_WINDOWS, margin=2, log=None):
    """lines: every record of the GPU run as VCF text lines.  Returns the per-window results."""
    base = g["inst_base"]
    n_inst = len(base) - 1
    ids = g["ids"].astype(np.int64)
    seglen = np.fromiter(map(len, g["seqs"]), dtype=np.int64, count=len(g["seqs"]))
    ref_steps = g["paths"][0][1]
    ref_cum = np.concatenate([[0], np.cumsum(seglen[ref_steps])])
    # index the GPU records by the smallest id of their ID column
    rec_ids = [[int(x) for x in re.findall(r"\d+", l.split("\t", 3)[2])] for l in lines]
    rec_lo = np.array([min(x) for x in rec_ids], dtype=np.int64)
    rec_hi = np.array([max(x) for x in rec_ids], dtype=np.int64)
    out = []
    for w in range(n_windows):
        k0 = int((w + 0.5) * n_inst / n_windows - width / 2)
        k0 = max(0, min(k0, n_inst - width)); k1 = k0 + width
        lo, hi = int(base[k0]), int(base[k1])
        sub = synth.window(g, lo, hi)
        d = tempfile.mkdtemp()
        gfa = os.path.join(d, "w.gfa")
        synth.write_gfa_fast(sub, gfa)
        t0 = time.time()
        r = subprocess.run([REF, "gfa2vcf", "-i", gfa, "-P", prefix, "-t", "1"], capture_output=True, text=True)
        t_ref = time.time() - t0
        subprocess.run(["rm", "-rf", d])
        assert r.returncode == 0, r.stderr[-500:]
        in_lo, in_hi = int(ids[base[k0 + margin]]), int(ids[base[k1 - margin]])      # stay clear of the two cuts
        first = np.nonzero((ref_steps >= lo) & (ref_steps < hi))[0]
        shift = int(ref_cum[first[0]]) if len(first) else 0

        def canon(line, add):
            f = line.rstrip("\n").split("\t")
            f[1] = str(int(f[1]) + add)
            return "\t".join(f)
        sel = np.nonzero((rec_lo >= in_lo) & (rec_hi < in_hi))[0]
        got = sorted(canon(lines[i], 0) for i in sel)
        want = []
        for l in r.stdout.splitlines():
            if l.startswith("#"):
                continue
            x = [int(v) for v in re.findall(r"\d+", l.split("\t", 3)[2])]
            if min(x) >= in_lo and max(x) < in_hi:
                want.append(canon(l, shift))
        want.sort()
        ok = got == want and len(want) > 0
        res = {"window": w, "instances": [k0, k1], "vertices": hi - lo, "records_compared": len(want), "subr": sum("VARTYPE=SUBR" in l for l in want),
               "tangled": sum("TANGLED=T" in l for l in want), "bit_exact": ok, "reference_s": round(t_ref, 2)}
        if not ok:
            gs, ws = set(got), set(want)
            res["only_gpu"] = sorted(gs - ws)[:3]; res["only_ref"] = sorted(ws - gs)[:3]
        out.append(res)
        if log:
            log(res)
    return out


def main():
    out = {}
    for cfg in sys.argv[1:]:
        t0 = time.time()
        g, prefix, width = build(cfg)
        V, E = len(g["ids"]), len(g["v1"])
        S = sum(len(p[1]) for p in g["paths"])
        t_gen = time.time() - t0
        print(cfg, f"generated V={V} E={E} S={S} in {t_gen:.1f}s", flush=True)
        pv = api.Povu(device=0) if not os.environ.get("PV_EMU_LIB") else api.Povu(lib_path=os.environ["PV_EMU_LIB"])
        t0 = time.time()
        s = api.Session.from_graph(pv, g)
        t_sess = time.time() - t0
        walls = []
        for _ in range(int(os.environ.get("SCALE_ITERS", "4"))):           # the page-locked result vectors and the device pool reach their steady state after 3 calls
            tw = time.time()
            ms = s.hot_path([prefix])
            walls.append(round((time.time() - tw) * 1e3, 2))
        prof_top = None
        if os.environ.get("SCALE_PROF", "1") != "0":
            pv.set_profiling(True)
            s.hot_path([prefix])
            prof = pv.profile()
            pv.set_profiling(False)
            tot = sum(k["ms"] for k in prof)
            prof_top = [{"name": k["name"], "launches": k["launches"], "ms": round(k["ms"], 3), "share": round(k["ms"] / tot, 3),
                         "GBps": round(k["bytes"] / (k["ms"] * 1e-3) / 1e9, 1) if k["ms"] > 0 else None} for k in sorted(prof, key=lambda k: -k["ms"])[:16]]
        st = s.call_stats()
        nrec = int(s.lib.pvgpu_session_record_count(s.h)); nfl = int(s.lib.pvgpu_session_flubble_count(s.h))
        assert st["hard_errors"] == 0
        t0 = time.time()
        lines, _ = s.record_lines()
        t_text = time.time() - t0
        print(cfg, f"hot path {ms} ms, {nrec} records, stats {st}", flush=True)
        wins = check_windows(g, lines, prefix, width, log=lambda r: print(cfg, "window", json.dumps(r), flush=True))
        ok = all(w["bit_exact"] for w in wins)
        out[cfg] = {"V": V, "E": E, "paths": len(g["paths"]), "steps": S, "flubbles": nfl, "records": nrec, "ms_upload_dec_index_call": ms, "host_wall_ms_per_call": walls, "stats": st,
                    "records_subr": sum("VARTYPE=SUBR" in l for l in lines), "records_tangled": sum("TANGLED=T" in l for l in lines),
                    "window_check": {"windows": wins, "bit_exact": ok, "records_compared": sum(w["records_compared"] for w in wins),
                                     "subr_compared": sum(w["subr"] for w in wins), "tangled_compared": sum(w["tangled"] for w in wins)},
                    "edges_per_s": E / ((ms[1] + ms[2] + ms[3]) * 1e-3), "records_per_s": nrec / ((ms[2] + ms[3]) * 1e-3),
                    "top_kernels": prof_top, "gen_s": t_gen, "session_s": t_sess, "text_s": t_text, "peak_device_GB": pv.peak_bytes / 1e9}
        print(cfg, json.dumps(out[cfg]), flush=True)
        os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
        json.dump(out, open(os.path.join(ROOT, "gpurun_out", "scale_run.json"), "w"), indent=1)
        assert ok, "a window differs from the reference"
        s.close(); pv.close()


if __name__ == "__main__":
    main()
