#!/usr/bin/env python
"""Per-kernel CUDA-event profile of one hot-path step on cfg2 (development aid): every kernel tag with launches,
ms and GB/s, the sum of kernel time and the un-profiled step time next to it (the difference = launch gaps + host)."""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
from povu_b200 import api, synth  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 400_000
paths = int(sys.argv[2]) if len(sys.argv) > 2 else 16
pv = api.Povu(device=0)
g = synth.snp_indel_chain(n, n_paths=paths, seed=42, with_seqs=True)
sess = api.Session.from_graph(pv, g, with_labels=True)
for _ in range(3):
    sess.hot_path(["S001#1"], keep_resident=True)
torch.cuda.synchronize()
s0 = pv.host_syncs
t0 = time.perf_counter()
for _ in range(5):
    parts = sess.hot_path(["S001#1"], keep_resident=True)
torch.cuda.synchronize()
wall = (time.perf_counter() - t0) / 5 * 1e3
print(f"host waits on the stream per step: {(pv.host_syncs - s0) / 5:.0f}")
l0 = pv.launches
pv.set_profiling(True)
sess.hot_path(["S001#1"], keep_resident=True)
prof = pv.profile()
pv.set_profiling(False)
tot = sum(k["ms"] for k in prof)
print(f"step wall {wall:.3f} ms (parts {['%.2f' % p for p in parts]}), kernel sum {tot:.3f} ms, launches {pv.launches - l0}")
for k in sorted(prof, key=lambda k: -k["ms"]):
    gbps = k["bytes"] / (k["ms"] * 1e-3) / 1e9 if k["ms"] > 0 else 0
    print(f"{k['name']:<18} x{k['launches']:<4} {k['ms']*1e3:9.1f} us  {100*k['ms']/tot:5.1f}%  {gbps:8.1f} GB/s")
