#!/usr/bin/env python
"""cfg4 AT ITS STATED SIZE (`wg_100M_25c` -- 25 weakly connected components with sizes proportional to GRCh38
chr1-22, X, Y, M, ~100 M segments / ~140 M links in total, 90 haplotype paths per component) on 1 / 2 / 4 / 8 B200.

The input arrives the way whole-genome pangenomes are distributed: ONE GRAPH PER CHROMOSOME (component).  Every rank knows the
25 component weights (here: from the generator's parameters; for files: their sizes), runs the same greedy LPT over them --
the partition rule of pvgpu_session_reshard, SURVEY 8e -- and ingests only the components it owns; each owned component is one
session (one decompose -> index -> call per step; no component holds more than 2^32 path steps, the whole graph does).  No
exchange while computing; the per-rank record counts are summed at the end (the packed-record NCCL gather is bench.py's
strong-scaling leg; here the records stay where the VCF writer of each rank would consume them).

Timing: per owned component W warm-up steps, then K steps between two CUDA events on the context's stream; a rank's step time is
the sum over its components, the job's step time the MAX over ranks (all-reduced).  Parity at this size: windows of the rank's
smallest owned component are cut out with their original ids, run through the UNMODIFIED reference (oracle/_ref/povu_ref) and
compared byte for byte with the records the device produced for the same regions (POS shifted by the window's offset on the
reference walk), every record type included.

usage:  python scripts/wg_full.py [--instances N] [--paths 90] [--steps 2] [--warmup 1] [--windows 2]
        python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 scripts/wg_full.py
result: gpurun_out/wg_full_n<N>.json (rank 0)"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402  (GRCH38 weights)

def templates():
    from povu_b200 import synth
    return [synth.template_from(synth.snp_indel_chain(6, n_paths=6, seed=100 + k)) for k in range(18)] + synth.cfg3_templates(seed=7, n_nested=2, tower_depth=0)


def lpt(weights, world):
    """greedy longest-processing-time: heaviest component first onto the lightest rank (ties: lowest rank) -- pv_shard.cuh's rule"""
    load = [0.0] * world
    owner = [0] * len(weights)
    for c in sorted(range(len(weights)), key=lambda c: (-weights[c], c)):
        r = min(range(world), key=lambda r: (load[r], r))
        owner[c] = r; load[r] += weights[c]
    return owner, load


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--instances", type=int, default=2_100_000)      # ~47.6 segments per instance -> ~100 M segments
    ap.add_argument("--paths", type=int, default=90)
    ap.add_argument("--steps", type=int, default=2)
    ap.add_argument("--warmup", type=int, default=1)
    ap.add_argument("--windows", type=int, default=2)
    ap.add_argument("--only", default="", help="comma-separated component numbers (diagnostics)")
    ap.add_argument("--profile", action="store_true", help="one extra step per component with per-kernel CUDA-event timing (top kernels to stderr)")
    ap.add_argument("--verbose", action="store_true", help="per step: the phases' wall ms and the DFS round count")
    ap.add_argument("--emu", action="store_true", help="host rehearsal library (logic check in the GPU-less container)")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1")); local = int(os.environ.get("LOCAL_RANK", "0"))
    from povu_b200 import api
    torch = dist = stream = None
    if not args.emu:
        import torch
        import torch.distributed as dist
        torch.cuda.set_device(local)
        if world > 1:
            dist.init_process_group("nccl", device_id=torch.device("cuda", local))
        stream = torch.cuda.Stream()
        pv = api.Povu(device=local, stream=stream.cuda_stream)
    else:
        pv = api.Povu(lib_path=os.path.join(ROOT, "tests", "emu", "libpovu_emu.so"))
    out = run_full(args, pv, stream, rank, world, torch, dist)
    if rank == 0:
        os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
        with open(os.path.join(ROOT, "gpurun_out", f"wg_full_n{world}{'_emu' if args.emu else ''}.json"), "w") as fh:
            json.dump(out, fh, indent=1)
        print(json.dumps({k: v for k, v in out.items() if k != "per_component"}))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    assert out is None or out["window_check"]["bit_exact"], "a window differs from the reference"


def run_full(args, pv, stream, rank, world, torch, dist):
    """args: instances, paths, steps, warmup, windows, only, profile, verbose, emu.  Returns the result dict on rank 0 (None elsewhere)."""
    from povu_b200 import api, synth
    w = np.asarray(bench.GRCH38, dtype=np.float64)
    cuts = np.round(np.cumsum(w) / w.sum() * args.instances).astype(np.int64); cuts[-1] = args.instances
    n_inst = np.diff(np.concatenate([[0], cuts]))
    owner, load = lpt([float(x) for x in n_inst], world)
    mine = [c for c in range(len(n_inst)) if owner[c] == rank and n_inst[c] > 0]
    if args.only:
        mine = [c for c in mine if str(c + 1) in args.only.split(",")]
    T = templates()
    tot = {"V": 0, "E": 0, "S": 0, "flubbles": 0, "records": 0, "ms": 0.0, "gen_s": 0.0, "ingest_s": 0.0, "e2e_ms": 0.0, "h2d_bytes": 0, "packed_bytes": 0}
    packed = []      # per owned component: its packed binary records in HBM (what this rank contributes to the final gather)
    per_comp, wins = [], []
    big = [c for c in mine if n_inst[c] >= 480]      # (chrM holds a dozen instances: no room for a window)
    smallest = min(big, key=lambda c: n_inst[c]) if big else None
    # every rank builds all the components it owns first and waits for the others: nobody times kernels while a neighbour's
    # generator keeps the host cores busy
    flats = {}
    gen_s = {}
    for c in mine:
        t0 = time.time()
        flats[c] = synth.tiled_flat(T, int(n_inst[c]), args.paths, seed=1000 + c, chrom=f"chr{c + 1}")
        gen_s[c] = time.time() - t0
    # ... and ingests them (one session per component: host arrays copied into the session, links / walks / segment lengths to HBM
    # on first use), so that no rank's timed steps run next to a neighbour's ingest
    sessions = {}
    ing_s = {}
    for c in mine:
        t1 = time.time()
        sessions[c] = api.Session.from_flat(pv, flats[c], with_labels=(args.windows > 0 and c == smallest))
        ing_s[c] = time.time() - t1
    if world > 1:
        dist.barrier()
    for c in mine:
        f = flats.pop(c)
        t2 = time.time(); t1 = t2 - ing_s[c]; t0 = t1 - gen_s[c]
        check = args.windows > 0 and c == smallest
        sess = sessions.pop(c)
        V, E, S = len(f["ids"]), len(f["v1"]), int(f["step_off"][-1])
        prefix = "S001#1"
        if args.emu:
            t = time.time(); sess.hot_path([prefix]); ms = (time.time() - t) * 1e3
        else:
            with torch.cuda.stream(stream):
                for _ in range(args.warmup):
                    sess.hot_path([prefix])
                torch.cuda.synchronize()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(stream)
                for _ in range(args.steps):
                    parts = sess.hot_path([prefix])
                    if args.verbose:
                        print(f"[rank {rank}] component {c + 1}: upload / decompose / index / call = {[round(x, 2) for x in parts]} ms, host waits so far {pv.host_syncs}, launches {pv.launches}", file=sys.stderr, flush=True)
                e1.record(stream)
                torch.cuda.synchronize()
                ms = e0.elapsed_time(e1) / args.steps
        if args.profile and not args.emu:
            pv.set_profiling(True)
            sess.hot_path([prefix])
            prof = pv.profile()
            pv.set_profiling(False)
            tk = sum(k["ms"] for k in prof)
            print(f"[rank {rank}] component {c + 1}: kernel sum {tk:.2f} ms, {sum(k['launches'] for k in prof)} launches", file=sys.stderr)
            for k in sorted(prof, key=lambda k: -k["ms"])[:28]:
                print(f"    {k['name']:<18} x{k['launches']:<4} {k['ms']:8.3f} ms {100 * k['ms'] / tk:5.1f}%  {k['bytes'] / (k['ms'] * 1e-3) / 1e9 if k['ms'] > 0 else 0:8.1f} GB/s", file=sys.stderr)
        # end to end for this component: raw arrays host -> HBM inside the step (links first, walks behind decompose), decompose,
        # index, call, the binary records packed into one stream in HBM (host wall clock around a synchronised step)
        e2e = 0.0
        if not args.emu:
            with torch.cuda.stream(stream):
                sess.hot_path([prefix], keep_resident=False)      # (warm: page-locks the host arrays once)
                torch.cuda.synchronize()
                reps = max(1, min(2, args.steps))
                tw = time.perf_counter()
                for _ in range(reps):
                    sess.hot_path([prefix], keep_resident=False)
                    need = sess.records_packed_bytes()
                    buf = torch.empty(need, dtype=torch.uint8, device="cuda")
                    sess.records_pack(buf.data_ptr(), need)
                torch.cuda.synchronize()
                e2e = (time.perf_counter() - tw) * 1e3 / reps
                packed.append(buf)
        nrec = int(sess.lib.pvgpu_session_record_count(sess.h)); nfl = int(sess.lib.pvgpu_session_flubble_count(sess.h))
        st = sess.call_stats()
        assert st["hard_errors"] == 0 and (nrec > 0 or n_inst[c] < 8)
        if check:
            lines, _ = sess.record_lines()
            import wg_windows      # (the only place the compiled reference is run: never reached from bench.py, which passes windows = 0)
            wins = wg_windows.check_windows(f, lines, prefix, args.windows)
            for x in wins:
                x["component"] = c + 1
            del lines
        sess.close()
        per_comp.append({"component": c + 1, "segments": V, "links": E, "path_steps": S, "flubbles": nfl, "records": nrec, "ms_per_step": round(ms, 3)})
        for k, v in (("V", V), ("E", E), ("S", S), ("flubbles", nfl), ("records", nrec), ("ms", ms), ("gen_s", t1 - t0), ("ingest_s", t2 - t1), ("e2e_ms", e2e),
                     ("h2d_bytes", 8 * E + 4 * S + 4 * V), ("packed_bytes", int(packed[-1].numel()) if packed else 0)):
            tot[k] += v
        del f
        print(f"[rank {rank}] component {c + 1}: V={V} E={E} S={S} {ms:.1f} ms/step, {nrec} records (generated in {t1 - t0:.1f} s, ingest {t2 - t1:.1f} s)", file=sys.stderr, flush=True)
    # the final gather: every rank's packed records to rank 0 over NCCL (sizes first), timed on the host around a synchronise
    gather_ms = 0.0
    if not args.emu:
        mine_bytes = torch.cat(packed) if packed else torch.empty(0, dtype=torch.uint8, device="cuda")
        del packed
        if world > 1:
            for rep in range(2):      # (first pass: NCCL sets its channels up)
                dist.barrier(); torch.cuda.synchronize()
                tw = time.perf_counter()
                sizes = [torch.zeros(1, dtype=torch.int64, device="cuda") for _ in range(world)]
                dist.all_gather(sizes, torch.tensor([mine_bytes.numel()], dtype=torch.int64, device="cuda"))
                mx = max(int(x.item()) for x in sizes)
                send = mine_bytes if mine_bytes.numel() == mx else torch.cat([mine_bytes, torch.empty(mx - mine_bytes.numel(), dtype=torch.uint8, device="cuda")])
                recv = [torch.empty(mx, dtype=torch.uint8, device="cuda") for _ in range(world)] if rank == 0 else None
                dist.gather(send, recv, dst=0)
                torch.cuda.synchronize()
                gather_ms = (time.perf_counter() - tw) * 1e3
                del recv, send
    tot["gather_ms"] = gather_ms
    res = {"rank": rank, "components": [p["component"] for p in per_comp], "per_component": per_comp, "windows": wins, **tot, "peak_device_GB": pv.peak_bytes / 1e9}
    if world > 1:
        allres = [None] * world
        dist.all_gather_object(allres, res)
    else:
        allres = [res]
    if rank != 0:
        return None
    if True:
        job_ms = max(r["ms"] for r in allres)
        E = sum(r["E"] for r in allres); V = sum(r["V"] for r in allres); S = sum(r["S"] for r in allres)
        rec = sum(r["records"] for r in allres)
        allw = [x for r in allres for x in r["windows"]]
        out = {"workload": f"cfg4 wg_100M_25c at its stated size: 25 components ~ GRCh38 chromosome sizes, {args.instances} template instances, {args.paths} paths per component; one graph per component, LPT by component weight, one session per owned component",
               "n_gpus": world, "segments": V, "links": E, "path_steps": S, "flubbles": sum(r["flubbles"] for r in allres), "records": rec,
               "ms_per_step": job_ms, "edges_per_s": E / (job_ms * 1e-3), "records_per_s": rec / (job_ms * 1e-3), "steps": args.steps, "warmup": args.warmup,
               "rank_ms": [round(r["ms"], 2) for r in allres], "lpt_load_max_over_mean": max(load) / (sum(load) / len(load)),
               "rank_components": [r["components"] for r in allres], "peak_device_GB_max": max(r["peak_device_GB"] for r in allres),
               "e2e": {"ms_per_step": max(r["e2e_ms"] for r in allres) + max(r["gather_ms"] for r in allres), "edges_per_s": E / (max(1e-9, max(r["e2e_ms"] for r in allres) + max(r["gather_ms"] for r in allres)) * 1e-3),
                       "rank_hot_path_from_host_ms": [round(r["e2e_ms"], 2) for r in allres], "nccl_gather_ms": round(max(r["gather_ms"] for r in allres), 2),
                       "h2d_bytes_per_step_all_ranks": sum(r["h2d_bytes"] for r in allres), "gathered_record_bytes": sum(r["packed_bytes"] for r in allres),
                       "note": "per owned component: links + walks + segment lengths host -> HBM inside the step, decompose, index, call, binary records packed in HBM; then ONE gather of all ranks' record streams on rank 0 over NCCL.  Host wall clock, max over ranks"},
               "generation_s_max": max(r["gen_s"] for r in allres), "ingest_s_max": max(r["ingest_s"] for r in allres),
               "window_check": {"windows": allw, "bit_exact": all(x["bit_exact"] for x in allw), "records_compared": sum(x["records_compared"] for x in allw)},
               "per_component": sorted((p for r in allres for p in r["per_component"]), key=lambda p: p["component"])}
        return out


if __name__ == "__main__":
    main()
