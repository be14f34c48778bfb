#!/usr/bin/env python
"""Where the time of pvgpu_session_reshard goes (PV_TRACE=1 prints the phases): one process plays rank 0 of `world` ranks
on the whole-genome-shaped graph of bench.py's strong-scaling leg."""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ.setdefault("PV_TRACE", "1")
import bench  # noqa: E402
from povu_b200 import api  # noqa: E402

world = int(sys.argv[1]) if len(sys.argv) > 1 else 2
inst = int(sys.argv[2]) if len(sys.argv) > 2 else 300_000
pv = api.Povu(device=0)
g = bench.wg_graph(inst, 24)
sess = api.Session.from_graph(pv, g, with_labels=False)
del g
for it in range(4):
    for from_host in (True,):
        t0 = time.perf_counter()
        n, ms = sess.reshard(0, world, from_host=from_host)
        print(f"reshard(0, {world}, from_host={from_host}): {ms:.2f} ms inside, {(time.perf_counter() - t0) * 1e3:.2f} ms wall, {n} components", file=sys.stderr, flush=True)
