#!/usr/bin/env python
"""GFA text -> graph arrays: device parser (pv_gfa.cuh) vs the host loader, on the cfg2 GFA text (development aid)."""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from povu_b200 import api, synth  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 400_000
pv = api.Povu(device=0)
g = synth.snp_indel_chain(n, n_paths=16, seed=42, with_seqs=True)
fp = "/tmp/gfa_parse_bench.gfa"
synth.write_gfa_fast(g, fp)
open(fp, "rb").read()          # page cache
for host in (False, False, False, True, False):
    if os.environ.get("PV_PROF") and not host:
        pv.set_profiling(True)
    t0 = time.perf_counter()
    s = api.Session(pv, fp, host_parser=host)
    wall = (time.perf_counter() - t0) * 1e3
    st = s.gfa_stats()
    print(f"{'host loader ' if host else 'device parse'}: open {wall:8.2f} ms   file->HBM {st['read_ms']:7.2f} ms  parse {st['parse_ms']:7.2f} ms   {os.path.getsize(fp) / 1e6:.1f} MB, "
          f"{os.path.getsize(fp) / wall / 1e6:.2f} GB/s", flush=True)
    if os.environ.get("PV_PROF") and not host:
        prof = pv.profile()
        pv.set_profiling(False)
        for k in sorted(prof, key=lambda k: -k["ms"])[:14]:
            print(f"    {k['name']:<18} x{k['launches']:<3} {k['ms'] * 1e3:9.1f} us")
    s.close()
