#!/usr/bin/env python
"""Window parity for scripts/wg_full.py (TEST SIDE: runs the compiled reference oracle/_ref/povu_ref).  Kept out of wg_full.py
so that bench.py, which imports wg_full for its whole_genome leg, never loads anything that touches oracle/."""
import os
import re
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
REF = os.path.join(ROOT, "oracle", "_ref", "povu_ref")


def check_windows(f, lines, prefix, n_windows, width=120, margin=2):
    from povu_b200 import synth
    base = f["inst_base"]; n_inst = len(base) - 1
    ids = f["ids"].astype(np.int64)
    ref_steps = (f["steps"][int(f["step_off"][0]):int(f["step_off"][1])] >> 1).astype(np.int64)
    ref_cum = np.concatenate([[0], np.cumsum(f["seg_len"][ref_steps].astype(np.int64))])
    rec_ids = [[int(x) for x in re.findall(r"\d+", l.split("\t", 3)[2])] for l in lines]
    rec_lo = np.array([min(x) for x in rec_ids], dtype=np.int64); rec_hi = np.array([max(x) for x in rec_ids], dtype=np.int64)
    out = []
    for w in range(n_windows):
        k0 = int((w + 0.5) * n_inst / n_windows - width / 2)
        k0 = max(0, min(k0, n_inst - width)); k1 = k0 + width
        lo, hi = int(base[k0]), int(base[k1])
        sub = synth.window_flat(f, lo, hi)
        d = tempfile.mkdtemp()
        gfa = os.path.join(d, "w.gfa")
        synth.write_gfa_fast(sub, gfa)
        t0 = time.time()
        r = subprocess.run([REF, "gfa2vcf", "-i", gfa, "-P", prefix, "-t", "1"], capture_output=True, text=True)
        t_ref = time.time() - t0
        subprocess.run(["rm", "-rf", d])
        assert r.returncode == 0, r.stderr[-500:]
        in_lo, in_hi = int(ids[base[k0 + margin]]), int(ids[base[k1 - margin]])
        first = int(np.searchsorted(ref_steps, lo, side="left"))
        shift = int(ref_cum[first])

        def canon(line, add):
            x = line.rstrip("\n").split("\t"); x[1] = str(int(x[1]) + add); return "\t".join(x)
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
        out.append({"window": w, "instances": [k0, k1], "vertices": hi - lo, "records_compared": len(want), "bit_exact": got == want and len(want) > 0, "reference_s": round(t_ref, 2)})
    return out
