#!/usr/bin/env python
"""bench.py -- povu_b200 hot-path benchmark (contract: task statement / DESIGN.md "Measurement").

One "step" = one pass of the whole decompose -> call hot path over one synthetic pangenome graph of the
shape BASELINE.json quotes its metric on (configs[1], `snp_indel_1M`: V = 1,000,001 segments,
E = 1,400,000 links, one component, 16 haplotype paths, reference path selected with -P 'S001#1'):
  decompose (components, exact DFS tree, cycle classes, flubbles, PVST)  ->  device index of the paths
  (adjacency in L-line order, loci, vertex->steps lists)  ->  call (regions, vertex order, depth
  matrices, contexts, allele groups, genotypes -> binary VCF records on the host).
`value` = graph edges/s through that whole step with the raw arrays (edge ports, packed path steps,
segment lengths) resident in HBM; `parts` splits it into BASELINE.json's two figures (decompose edges/s,
call VCF records/s).  `e2e` = the same step through the C ABI with pinned HOST arrays (H2D of the raw
arrays and D2H of flubbles + records inside the timed region).  N > 1: one process per GPU, every rank
runs its own component of that shape (weak scaling, no data-path collective: components are independent).

`--impl reference` times the reference's own CPU implementation (oracle/_ref/povu_ref = the unmodified
reference sources compiled by oracle/Makefile) on a bounded sample of the same workload.
"""
import argparse
import json
import os
import re
import subprocess
import sys
import tempfile
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

CFG2_BUBBLES = 400_000
N_PATHS = 16
REF_PREFIX = "S001#1"
METRIC = "decompose_call_hot_path_edges_per_sec"


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="native", choices=["native", "reference"])
    ap.add_argument("--bubbles", type=int, default=CFG2_BUBBLES, help="bubbles in the cfg2-shaped chain (default = cfg2)")
    ap.add_argument("--paths", type=int, default=N_PATHS)
    ap.add_argument("--cpu-bubbles", type=int, default=0, help="bubbles of the CPU arm's graph (default 0 = the same graph as the GPU arm: same config)")
    ap.add_argument("--ref-budget-s", type=float, default=1500.0, help="reference arm: stop adding timed passes when the run would exceed this wall time (the line then says how many were timed)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-gfa-leg", action="store_true", help="skip the GFA text -> HBM leg (device parser) reported under gfa_text")
    ap.add_argument("--profile-steps", type=int, default=2)
    ap.add_argument("--no-strong", action="store_true", help="skip the strong-scaling leg (whole-genome-shaped graph sharded over the ranks)")
    ap.add_argument("--no-wg-full", action="store_true", help="skip the whole-genome leg at cfg4's stated size (scripts/wg_full.py: ~100 M segments, 90 paths per component; about a minute at N = 1)")
    ap.add_argument("--wg-full-instances", type=int, default=2_100_000, help="template instances of the full-size whole-genome graph (2.1 M = 100.4 M segments / 142 M links / 2.6 G path steps)")
    ap.add_argument("--wg-instances", type=int, default=300_000, help="template instances of the whole-genome-shaped graph (300 k = 14.3 M segments / 20.2 M links)")
    ap.add_argument("--wg-paths", type=int, default=24, help="haplotype paths per component of that graph")
    return ap.parse_args()


def workload_name(n, paths):
    return f"snp_indel_1M (cfg2), {paths} haplotype paths" if n == CFG2_BUBBLES else f"snp_indel_chain({n} bubbles), {paths} haplotype paths"


# ----------------------------------------------------------------------------------------------
# CPU reference leg
# ----------------------------------------------------------------------------------------------
_GFA_CACHE = {}


def cpu_gfa(n_bubbles, n_paths, seed=42):
    """The workload as GFA text for the CPU arm (written once per process: the generator and the 141 MB text are not the thing timed)."""
    key = (n_bubbles, n_paths, seed)
    if key not in _GFA_CACHE:
        from povu_b200 import synth
        g = synth.snp_indel_chain(n_bubbles, n_paths=n_paths, seed=seed, with_seqs=True)
        d = tempfile.mkdtemp(prefix="povu_bench_")
        gfa = os.path.join(d, "g.gfa")
        synth.write_gfa_fast(g, gfa)
        _GFA_CACHE[key] = (d, gfa, len(g["v1"]))
        import atexit
        atexit.register(lambda d=d: subprocess.run(["rm", "-rf", d]))
    return _GFA_CACHE[key]


def cpu_reference_run(n_bubbles, n_paths, seed=42):
    """Reference decompose + call on a cfg2-shaped GFA; compute time = the six decompose stage timers +
    (do_call wall - a stand-alone GFA load with the same configuration)."""
    d, gfa, E = cpu_gfa(n_bubbles, n_paths, seed)
    ref = os.path.join(ROOT, "oracle", "_ref", "povu_ref")
    port = os.path.join(ROOT, "oracle", "povu_oracle")
    env = dict(os.environ, POVU_STAGE_COST_TRACE="1")
    threads = str(min(os.cpu_count() or 1, 255))      # the reference parallelises decompose across components only (one here); call = producer + writer thread
    if os.path.exists(ref):
        t0 = time.perf_counter()
        with open(os.path.join(d, "out.vcf"), "w") as so:
            r = subprocess.run([ref, "timed-gfa2vcf", "-i", gfa, "-P", REF_PREFIX, "-t", threads], env=env, stdout=so, stderr=subprocess.PIPE, text=True)
        wall = time.perf_counter() - t0
        if r.returncode != 0:
            raise RuntimeError("CPU reference failed: " + r.stderr[-300:])
        stage = sum(int(x) for x in re.findall(r"povu-stage-cost contract=\S+ .*elapsed_ns=(\d+)", r.stderr)) / 1e9
        m = re.search(r"povu-ref-timing decompose_total_s=(\S+) call_load_s=(\S+) call_total_s=(\S+)", r.stderr)
        dec_total, call_load, call_total = (float(x) for x in m.groups())
        records = sum(1 for l in open(os.path.join(d, "out.vcf")) if not l.startswith("#"))
        t_dec = stage if stage > 0 else dec_total
        t_call = max(call_total - call_load, 1e-9)
        return {"edges": E, "records": records, "t_dec": t_dec, "t_call": t_call, "wall_s": wall, "kind": "reference", "cores": 2,
                "note": f"-t {threads}; decompose = sum of the reference's six stage timers (1 busy thread: it parallelises only across components, and this graph is one); call = do_call wall minus a stand-alone GFA load (producer thread + VCF text consumer thread = the 2 cores counted)"}
    t0 = time.perf_counter()
    r = subprocess.run([port, "gfa2vcf", "-i", gfa, "-P", REF_PREFIX], stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)
    wall = time.perf_counter() - t0
    if r.returncode != 0:
        raise RuntimeError("oracle port failed: " + r.stderr[-300:])
    records = sum(1 for l in r.stdout.splitlines() if not l.startswith("#"))
    return {"edges": E, "records": records, "t_dec": wall / 2, "t_call": wall / 2, "wall_s": wall, "kind": "port", "cores": 1, "note": "oracle port, whole-process wall incl. GFA parse"}


def cpu_fields(r, n, paths):
    t = r["t_dec"] + r["t_call"]
    return {"value": r["edges"] / t, "unit": "edges/s", "cores": r["cores"], "kind": r["kind"],
            "sample": f"{workload_name(n, paths)}: {r['edges']} edges, {r['records']} records, one pass; {r['note']}; decompose {r['t_dec']:.2f} s + call {r['t_call']:.2f} s (whole process wall {r['wall_s']:.2f} s incl. two GFA parses)",
            "decompose_edges_per_sec": r["edges"] / r["t_dec"], "call_records_per_sec": r["records"] / r["t_call"], "host_cores_available": os.cpu_count()}


def base_config(n, paths):
    """config keys shared by both arms (the driver compares them): the workload both arms run"""
    n_snp, n_ind = (n + 1) // 2, n // 2
    return {"workload": workload_name(n, paths), "segments": (n + 1) + 2 * n_snp + n_ind, "links": 4 * n_snp + 3 * n_ind, "paths": paths, "components_per_gpu": 1, "ref_prefix": REF_PREFIX}


def run_reference_arm(args):
    """The reference's own CPU implementation on the SAME graph as the GPU arm (full cfg2 by default: ~17 s per pass).  One
    untimed warm-up pass (page cache, binary load), then --steps timed passes; if the whole run would exceed --ref-budget-s
    the remaining passes are dropped and the line says how many were timed (never fewer than 3)."""
    if int(os.environ.get("RANK", "0")) != 0:
        return
    n = args.cpu_bubbles or args.bubbles
    t_start = time.perf_counter()
    cpu_gfa(n, args.paths)
    vals, times, last = [], [], None
    if args.warmup > 0:
        cpu_reference_run(n, args.paths)
    per_pass = None
    for i in range(args.steps):
        t0 = time.perf_counter()
        last = cpu_reference_run(n, args.paths)
        per_pass = time.perf_counter() - t0
        t = last["t_dec"] + last["t_call"]
        vals.append(last["edges"] / t); times.append(t)
        if len(vals) >= 3 and (time.perf_counter() - t_start) + per_pass > args.ref_budget_s:
            break
    v = len(vals) * last["edges"] / sum(times)
    cpu = cpu_fields(last, n, args.paths)
    cpu["value"] = v
    cfg = base_config(n, args.paths)
    cfg.update({"records": last["records"], "timed_passes": len(vals), "same_graph_as_gpu_arm": n == args.bubbles})
    line = {"impl": "reference", "metric": METRIC, "value": v, "unit": "edges/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": 1e3 * sum(times) / len(times), "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u32", "data": "synthetic",
            "config": cfg,
            "parts": {"decompose_edges_per_sec": cpu["decompose_edges_per_sec"], "call_records_per_sec": cpu["call_records_per_sec"], "records": last["records"]},
            "cpu_baseline": cpu, "e2e": {"value": v, "unit": "edges/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


# ----------------------------------------------------------------------------------------------
# clocks
# ----------------------------------------------------------------------------------------------
class ClockSampler:
    Q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        self.index = index
        self.rows = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "50"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.perf_counter(), [x.strip() for x in line.split(",")]))

    def mark_begin(self):
        self.t0 = time.perf_counter()

    def mark_end(self):
        self.t1 = time.perf_counter()

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], None, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        # the sampler runs from before the warm-up (nvidia-smi needs ~50 ms to deliver its first row); keep the rows that arrived
        # inside the timed region, widened by one sampling period on both sides
        t0, t1 = getattr(self, "t0", None), getattr(self, "t1", None)
        rows = [r for t, r in self.rows if t0 is None or t1 is None or (t0 - 0.06 <= t <= t1 + 0.06)]
        if not rows:
            rows = [r for _, r in self.rows]
        for r in rows:
            if len(r) < 7:
                continue
            try:
                sm.append(float(r[0])); mx = float(r[1])
            except ValueError:
                continue
            for k, nme in enumerate(names):
                if r[3 + k].lower().startswith("active"):
                    reasons.add(nme)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons), "samples": len(sm)}


# ----------------------------------------------------------------------------------------------
# strong scaling: ONE whole-genome-shaped graph, its components partitioned over the ranks (north_star / SURVEY.md 8e)
# ----------------------------------------------------------------------------------------------
GRCH38 = [248956422, 242193529, 198295559, 190214555, 181538259, 170805979, 159345973, 145138636, 138394717, 133797422, 135086622, 133275309,
          114364328, 107043718, 101991189, 90338345, 83257441, 80373285, 58617616, 64444167, 46709983, 50818468, 156040895, 57227415, 16569]


def wg_graph(instances, paths):
    """cfg4's shape: 25 weakly connected components with sizes proportional to GRCh38 chr1-22, X, Y, M; SNP / indel
    runs with ~10 % nested templates; `paths` haplotype paths per component.  Deterministic: every rank builds the same graph."""
    from povu_b200 import synth
    tpl = [synth.template_from(synth.snp_indel_chain(6, n_paths=6, seed=100 + k)) for k in range(18)] + synth.cfg3_templates(seed=7, n_nested=2, tower_depth=0)
    return synth.tiled(tpl, instances, paths, seed=11, n_components=25, comp_weights=GRCH38)


def run_strong(args, pv, stream, rank, world, torch, dist):
    """value: device-timed step on the resident shard (one decompose + index per step, then one call per owned component against
    that component's reference path -- a region is only called where EVERY selected reference runs, so the reference too needs one
    run per chromosome).  e2e: the same from the page-locked host arrays of the WHOLE graph: host -> HBM, partition on the device
    (pvgpu_session_reshard), hot path, the binary records packed in HBM and gathered to rank 0 over NCCL."""
    from povu_b200 import api
    t0 = time.perf_counter()
    g = wg_graph(args.wg_instances, args.wg_paths)
    V, E = len(g["ids"]), len(g["v1"])
    S = sum(len(p[1]) for p in g["paths"])
    gen_s = time.perf_counter() - t0
    sess = api.Session.from_graph(pv, g, with_labels=False)
    del g

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def owned_refs():
        ga = sess.graph_arrays(labels=False)
        return [n for k, n in enumerate(ga["path_names"]) if n.startswith("S001#1#") and ga["step_off"][k + 1] > ga["step_off"][k]]

    n_comp, shard_ms0 = sess.reshard(rank, world)
    v_, e_, h_, st_ = sess.sizes()
    own = owned_refs()
    buf = torch.empty(64 << 20, dtype=torch.uint8, device="cuda")

    def step(keep, pack):
        nonlocal buf
        nrec = used = 0
        for i, pfx in enumerate(own):
            if i == 0:
                sess.hot_path([pfx], keep_resident=keep)
            else:
                sess.call([pfx])
            nrec += int(sess.lib.pvgpu_session_record_count(sess.h))
            if pack:
                need = sess.records_packed_bytes()
                if used + need > buf.numel():
                    nb = torch.empty(max(2 * buf.numel(), used + need), dtype=torch.uint8, device="cuda")
                    nb[:used] = buf[:used]; buf = nb
                used += sess.records_pack(buf.data_ptr() + used, buf.numel() - used)
        return nrec, used

    with torch.cuda.stream(stream):
        for _ in range(args.warmup):
            step(True, False)
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(args.steps):
            nrec, _ = step(True, False)
        e1.record(stream)
        barrier()
        dev_ms = e0.elapsed_time(e1) / args.steps
        nfl = int(sess.lib.pvgpu_session_flubble_count(sess.h))

        # end to end: whole graph on the host -> partition -> hot path -> packed records -> gather on rank 0
        def e2e_step():
            t_a = time.perf_counter()
            t_sh = 0.0
            if world > 1:
                _, t_sh = sess.reshard(rank, world, from_host=True)
            t_b = time.perf_counter()
            nrec, used = step(world > 1, True) if world > 1 else step(False, True)
            torch.cuda.synchronize()
            t_c = time.perf_counter()
            sizes = torch.tensor([used], dtype=torch.int64, device="cuda")
            gathered = 0
            if world > 1:
                all_sizes = [torch.zeros_like(sizes) for _ in range(world)]
                dist.all_gather(all_sizes, sizes)
                mx = int(max(int(x.item()) for x in all_sizes))
                send = buf[:mx] if buf.numel() >= mx else torch.cat([buf, torch.empty(mx - buf.numel(), dtype=torch.uint8, device="cuda")])
                recv = [torch.empty(mx, dtype=torch.uint8, device="cuda") for _ in range(world)] if rank == 0 else None
                dist.gather(send, recv, dst=0)
                torch.cuda.synchronize()
                gathered = sum(int(x.item()) for x in all_sizes)
            else:
                gathered = used
            t_d = time.perf_counter()
            return (t_b - t_a) * 1e3, used, gathered, (t_c - t_b) * 1e3, (t_d - t_c) * 1e3
        for _ in range(min(2, args.warmup)):
            e2e_step()
        barrier()
        n_e2e = max(3, min(args.steps, 5))
        t0 = time.perf_counter()
        sh, hp, ga = [], [], []
        for _ in range(n_e2e):
            t_sh, used, gathered, t_hot, t_ga = e2e_step()
            sh.append(t_sh); hp.append(t_hot); ga.append(t_ga)
        barrier()
        e2e_ms = (time.perf_counter() - t0) * 1e3 / n_e2e

    t = torch.tensor([dev_ms, e2e_ms, sum(sh) / len(sh), sum(hp) / len(hp), sum(ga) / len(ga)], dtype=torch.float64, device="cuda")
    cnt = torch.tensor([nrec, nfl, e_], dtype=torch.float64, device="cuda")
    share = torch.tensor([float(e_ + st_ + v_)], dtype=torch.float64, device="cuda")
    shares = [torch.zeros_like(share) for _ in range(world)]
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(cnt, op=dist.ReduceOp.SUM)
        dist.all_gather(shares, share)
    else:
        shares = [share]
    sess.close()
    w = [float(x.item()) for x in shares]
    dev_ms, e2e_ms, shard_ms, hot_ms, gather_ms = t.tolist()
    return {"workload": f"wg_25c (cfg4 shape): 25 components ~ GRCh38 chromosome sizes, {args.wg_instances} template instances, {args.wg_paths} paths per component",
            "scaling": "strong", "n_gpus": world, "segments": V, "links": E, "path_steps": S, "components": n_comp,
            "value": E / (dev_ms * 1e-3), "unit": "edges/s", "ms_per_step": dev_ms, "steps": args.steps,
            "records_all_ranks": int(cnt[0].item()), "flubbles_all_ranks": int(cnt[1].item()), "links_all_ranks": int(cnt[2].item()),
            "lpt_load_max_over_mean": max(w) / (sum(w) / len(w)), "calls_per_step_rank0": len(own),
            "e2e": {"value": E / (e2e_ms * 1e-3), "unit": "edges/s", "ms_per_step": e2e_ms, "steps": n_e2e, "breakdown_ms_max_over_ranks": {"h2d_whole_graph_links_and_segments_plus_device_partition": shard_ms, "hot_path_calls_and_record_packing": hot_ms, "size_exchange_and_nccl_gather": gather_ms},
                    "h2d_bytes_per_step_per_rank": (8 * E + 8 * V + (4 * S) // world) if world > 1 else 8 * E + 4 * S + 4 * V, "gathered_record_bytes": int(gathered),
                    "note": "per step and rank: whole graph host -> HBM, device partition (labels, weights, compaction), decompose + index + one call per owned component, binary records packed in HBM; then all ranks' streams gathered on rank 0 over NCCL (N = 1: no partition, no gather).  Every rank needs ALL links to label the components, so that part of the ingest does not shrink with N; the walks (the bulk) travel only for the paths a rank owns"},
            "graph_generation_s": gen_s}


# ----------------------------------------------------------------------------------------------
# native arm
# ----------------------------------------------------------------------------------------------
def run_native(args):
    import torch
    import torch.distributed as dist

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- povu_b200 has no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    import __graft_entry__ as ge
    if rank == 0:
        ge.build_cuda()
    if world > 1:
        dist.barrier()
    from povu_b200 import api, synth

    stream = torch.cuda.Stream()
    pv = api.Povu(device=local, stream=stream.cuda_stream)

    # every rank gets its own component of the cfg2 shape (different seed => different alleles / paths)
    g = synth.snp_indel_chain(args.bubbles, n_paths=args.paths, seed=42 + rank, with_seqs=True)
    V, E = len(g["ids"]), len(g["v1"])
    sess = api.Session.from_graph(pv, g, with_labels=True)
    S = sess.n_steps
    del g
    prefixes = [REF_PREFIX]
    with torch.cuda.stream(stream):
        flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")   # > 126 MB L2

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    clocks = ClockSampler(local)
    if rank == 0:
        clocks.start()
    with torch.cuda.stream(stream):
        for _ in range(args.warmup):
            sess.hot_path(prefixes, keep_resident=True)
        barrier()
        clocks.mark_begin()
        launches0 = pv.launches
        syncs0 = pv.host_syncs
        evs, parts = [], []
        wall0 = time.perf_counter()
        for _ in range(args.steps):
            flush.fill_(1)                      # evict the previous step's working set from L2 (untimed)
            e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            parts.append(sess.hot_path(prefixes, keep_resident=True))
            e1.record(stream)
            evs.append((e0, e1))
        barrier()
        clocks.mark_end()
        wall = time.perf_counter() - wall0
        launches = pv.launches - launches0
        host_syncs = pv.host_syncs - syncs0
        dev_ms = sum(a.elapsed_time(b) for a, b in evs)
        clk = clocks.stop() if rank == 0 else None
        n_rec = int(sess.lib.pvgpu_session_record_count(sess.h))
        n_fl = int(sess.lib.pvgpu_session_flubble_count(sess.h))
        stats = sess.call_stats()

        # end-to-end through the C ABI from pinned HOST arrays (raw arrays dropped from the device every step)
        for _ in range(min(2, args.warmup)):
            sess.hot_path(prefixes, keep_resident=False)
        barrier()
        e2e_steps = max(3, min(args.steps, 10))
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            sess.hot_path(prefixes, keep_resident=False)
        barrier()
        e2e_s = time.perf_counter() - t0
        # VCF text of the records, written on the device (pv_text.cuh) and brought to the host; reported beside e2e
        devnull = os.path.join(tempfile.gettempdir(), f"povu_bench_{os.getpid()}.vcf")
        sess.vcf_write(devnull)                    # first call uploads the per-graph text tables
        t0 = time.perf_counter()
        text_bytes = sess.vcf_write("/dev/null")
        text_s = time.perf_counter() - t0
        os.unlink(devnull)

        # the steps either side of the hot path, on rank 0: GFA text parsed on the device (pv_gfa.cuh) and the
        # whole file -> file chain (open + decompose + index + call + VCF text) in one process, CUDA context already up
        gfa_leg = None
        if rank == 0 and not args.no_gfa_leg:
            gfa_fp = os.path.join(tempfile.gettempdir(), f"povu_bench_{os.getpid()}.gfa")
            synth.write_gfa_fast(synth.snp_indel_chain(args.bubbles, n_paths=args.paths, seed=42 + rank, with_seqs=True), gfa_fp)
            try:
                opens, chain = [], []
                for i in range(4):
                    t0 = time.perf_counter()
                    s2 = api.Session(pv, gfa_fp)
                    t1 = time.perf_counter()
                    s2.hot_path(prefixes, keep_resident=True)
                    s2.vcf_write("/dev/null")
                    t2 = time.perf_counter()
                    st = s2.gfa_stats()
                    s2.close()
                    if i:                                   # the first open allocates the page-locked staging pool
                        opens.append(((t1 - t0) * 1e3, st["read_ms"], st["parse_ms"]))
                        chain.append((t2 - t0) * 1e3)
                t0 = time.perf_counter()
                s2 = api.Session(pv, gfa_fp, host_parser=True)
                host_ms = (time.perf_counter() - t0) * 1e3
                s2.close()
                best = min(opens)
                gfa_leg = {"text_bytes": st["text_bytes"], "open_ms": best[0], "file_to_hbm_ms": best[1], "device_parse_ms": best[2],
                           "text_GBps": st["text_bytes"] / best[0] / 1e6, "host_loader_ms": host_ms,
                           "gfa_file_to_vcf_file_ms": min(chain),
                           "note": "pvgpu_session_open on the cfg2 GFA text in the page cache: threaded pread -> page-locked staging -> HBM, parsed by the G_* kernels; "
                                   "host_loader_ms = the sequential host loader on the same file; gfa_file_to_vcf_file_ms = open + decompose + index + call + device VCF text + write(2)"}
            finally:
                os.unlink(gfa_fp)

        prof = None
        if rank == 0 and args.profile_steps > 0:
            pv.set_profiling(True)
            for _ in range(args.profile_steps):
                flush.fill_(1)
                sess.hot_path(prefixes, keep_resident=True)
            prof = pv.profile()
            pv.set_profiling(False)

    strong = None
    if not args.no_strong:
        sess.close()
        try:
            strong = run_strong(args, pv, stream, rank, world, torch, dist)
        except Exception as ex:      # reported, never fatal for the headline line
            strong = {"error": str(ex)[:300]}
    # cfg4 at its stated size, one graph per component (scripts/wg_full.py): every rank generates and ingests only the components
    # the LPT rule gives it; device-timed step = max over ranks of the sum over its components; e2e = the same from host arrays +
    # one NCCL gather of the packed records.  (The byte-for-byte windows against the reference are that script's own job: bench.py
    # may run the oracle only as the CPU baseline.)
    whole = None
    if not args.no_wg_full:
        if args.no_strong:
            sess.close()
        try:
            import types
            sys.path.insert(0, os.path.join(ROOT, "scripts"))
            import wg_full
            wa = types.SimpleNamespace(instances=args.wg_full_instances, paths=90, steps=2, warmup=3, windows=0, only="", profile=False, verbose=False, emu=False)
            whole = wg_full.run_full(wa, pv, stream, rank, world, torch, dist)
            if whole:
                whole.pop("window_check", None)
        except Exception as ex:      # reported, never fatal for the headline line
            whole = {"error": str(ex)[:300]}
    t_dev = torch.tensor([dev_ms, e2e_s * 1e3], dtype=torch.float64, device="cuda")
    cnt = torch.tensor([n_rec, n_fl], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t_dev, op=dist.ReduceOp.MAX)
        dist.all_reduce(cnt, op=dist.ReduceOp.SUM)     # the only exchange: totals of the per-rank record buffers
    dev_ms_max, e2e_ms_max = t_dev.tolist()
    tot_rec, tot_fl = (int(x) for x in cnt.tolist())

    if rank == 0:
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        peak = float(peaks.get("hbm_gbs", 6650.0))
        peak_src = "measured (MEASURED_PEAKS.json hbm_gbs)" if "hbm_gbs" in peaks else "fallback 6650 GB/s (B200_PROFILING.md)"
        if not prof:
            prof = [{"name": "unprofiled", "ms": 0.0, "bytes": 0, "launches": 1, "items": 0}]
            args.profile_steps = 1
        total_ms = sum(k["ms"] for k in prof) or 1.0
        top = max(prof, key=lambda k: k["ms"])
        ach = top["bytes"] / (top["ms"] * 1e-3) / 1e9 if top["ms"] > 0 else 0.0
        kernels = sorted(prof, key=lambda k: -k["ms"])[:10]
        ms_per_step = dev_ms_max / args.steps
        # compulsory bytes of the whole step: decompose 16E + 12V + 16F ; call 8*S + 16F + ~(48 + H + 16) per record
        comp_dec = 16 * E + 12 * V + 16 * n_fl
        comp_call = 8 * S + 16 * n_fl + n_rec * (48 + args.paths + 16)
        traffic = None
        try:
            traffic = json.load(open(os.path.join(ROOT, "profiles", "traffic.json"))).get(top["name"], {}).get("dram_bytes_per_launch")
        except Exception:
            pass
        roofline = {"bound": "hbm", "kernel": top["name"], "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak,
                    "traffic": traffic, "peak_source": peak_src, "kernel_share_of_step": top["ms"] / total_ms,
                    "avg_launch_us": 1e3 * top["ms"] / top["launches"], "launches_per_step": top["launches"] / args.profile_steps,
                    "algorithmic_bytes_per_launch": top["bytes"] / top["launches"],
                    "pipeline_compulsory_bytes": {"decompose": comp_dec, "call": comp_call},
                    "pipeline_frac_of_peak": (comp_dec + comp_call) / (ms_per_step * 1e-3) / 1e9 / peak,
                    "top_kernels": [{"name": k["name"], "ms_per_step": k["ms"] / args.profile_steps, "launches_per_step": k["launches"] / args.profile_steps,
                                     "GBps": (k["bytes"] / (k["ms"] * 1e-3) / 1e9) if k["ms"] > 0 else None} for k in kernels]}
        cpu = None
        if not args.no_cpu_baseline and world == 1:
            try:
                nb = args.cpu_bubbles or args.bubbles
                cpu = cpu_fields(cpu_reference_run(nb, args.paths), nb, args.paths)
                cpu["same_config"] = nb == args.bubbles
            except Exception as ex:   # the baseline is reported, never required
                cpu = {"value": None, "unit": "edges/s", "cores": 0, "kind": "unavailable", "sample": str(ex)[:200]}
        pm = [sum(p[i] for p in parts) / len(parts) for i in range(4)]
        d2h_bytes = n_fl * 22 + n_rec * (9 * 4 + 8) + n_rec * 2 * 16 + n_rec * args.paths * 2
        line = {"metric": METRIC, "value": world * E / (ms_per_step * 1e-3), "unit": "edges/s", "n_gpus": world,
                "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": "u32", "data": "synthetic",
                "config": {**base_config(args.bubbles, args.paths), "path_steps": S, "flubbles": n_fl, "records": n_rec,
                           "l2": "256 MiB flush write between timed steps (untimed); per-step CUDA events summed",
                           "sharding": "one independent component per rank, no collective on the data path",
                           "outputs": "value: the flubble table reaches the host (copy overlapped), the binary records stay in HBM where the device VCF writer reads them; "
                                      "e2e: inputs from page-locked host arrays (walks uploaded behind decompose) and the binary records copied to the host as well"},
                "parts": {"decompose_edges_per_sec": world * E / (pm[1] * 1e-3), "call_records_per_sec": tot_rec / ((pm[2] + pm[3]) * 1e-3),
                          "ms": {"decompose": pm[1], "path_index_build": pm[2], "call": pm[3]}, "records_all_ranks": tot_rec, "flubbles_all_ranks": tot_fl, "call_stats": stats},
                "e2e": {"value": world * E / (e2e_ms_max / e2e_steps * 1e-3), "unit": "edges/s", "h2d_bytes_per_step": 8 * E + 4 * S + 4 * V, "d2h_bytes_per_step": d2h_bytes,
                        "ms_per_step": e2e_ms_max / e2e_steps, "steps": e2e_steps,
                        "vcf_text": {"bytes": text_bytes, "ms": text_s * 1e3, "note": "record lines written by the device text kernels (X_len, X_write), D2H into a page-locked buffer, write(2) to /dev/null; wall time of one pvgpu_session_vcf_write call, not inside e2e (SURVEY.md 8f2)"}},
                "gfa_text": gfa_leg, "gpu_launches": int(launches), "host_syncs_per_step": host_syncs / args.steps, "wall_s_timed_region": wall,
                "clocks": clk, "roofline": roofline, "cpu_baseline": cpu, "strong_scaling": strong, "whole_genome": whole, "peak_device_bytes": pv.peak_bytes}
        print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main():
    args = parse_args()
    if args.impl == "reference":
        run_reference_arm(args)
    else:
        run_native(args)


if __name__ == "__main__":
    main()
