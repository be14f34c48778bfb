"""Coarse phase trace of one hot-path step (PV_TRACE=1 synchronises between phases)."""
import os
import sys
os.environ["PV_TRACE"] = "1"
sys.path.insert(0, os.getcwd())
from povu_b200 import api, synth
pv = api.Povu(device=0)
g = synth.snp_indel_chain(400000, n_paths=16, seed=42)
s = api.Session.from_graph(pv, g)
for i in range(4):
    print("=== step", i, file=sys.stderr, flush=True)
    s.hot_path(["S001#1"])
