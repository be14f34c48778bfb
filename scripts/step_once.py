#!/usr/bin/env python
"""N hot-path steps on cfg2 and nothing else (the command ncu wraps: launch lists, --set full captures)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from povu_b200 import api, synth  # noqa: E402

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 2
n = int(sys.argv[2]) if len(sys.argv) > 2 else 400_000
paths = int(sys.argv[3]) if len(sys.argv) > 3 else 16
pv = api.Povu(device=0)
sess = api.Session.from_graph(pv, synth.snp_indel_chain(n, n_paths=paths, seed=42, with_seqs=True), with_labels=True)
for _ in range(steps):
    sess.hot_path(["S001#1"], keep_resident=True)
print("launches", pv.launches, "per step", pv.launches / steps)
