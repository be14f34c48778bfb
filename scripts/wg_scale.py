#!/usr/bin/env python
"""Whole-genome-shaped STRONG scaling: one graph of 25 weakly connected components with sizes
proportional to GRCh38 chr1-22, X, Y, M; every rank loads the same graph, keeps its LPT share of the components
(pvgpu_session_shard) and runs decompose -> index -> call on it.  No collective on the data path; the ranks only exchange
their timings.  value = edges of the WHOLE graph / max-over-ranks step time.
    python scripts/wg_scale.py [--instances N] [--paths H]                              (1 GPU)
    python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 scripts/wg_scale.py ...
The graph is scaled down from cfg4 (100 M segments, 90 paths) by --instances so that the host-side generator stays in
seconds; the shape (component size distribution, bubble mix: SNP / indel chains with ~10 % nested templates) is cfg4's."""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np   # noqa: E402
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402
from povu_b200 import api, synth  # noqa: E402

GRCH38 = [248956422, 242193529, 198295559, 190214555, 181538259, 170805979, 159345973, 145138636, 138394717, 133797422, 135086622, 133275309,
          114364328, 107043718, 101991189, 90338345, 83257441, 80373285, 58617616, 64444167, 46709983, 50818468, 156040895, 57227415, 16569]


def templates():
    simple = [synth.template_from(synth.snp_indel_chain(6, n_paths=6, seed=100 + k)) for k in range(18)]      # SNP / indel runs
    nested = synth.cfg3_templates(seed=7, n_nested=2, tower_depth=0)                                          # ~10 % nested
    return simple + nested


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--instances", type=int, default=300_000)
    ap.add_argument("--paths", type=int, default=24)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    a = ap.parse_args()
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    g = synth.tiled(templates(), a.instances, a.paths, seed=11, n_components=25, comp_weights=GRCH38)
    V, E = len(g["ids"]), len(g["v1"])
    S = sum(len(p[1]) for p in g["paths"])
    stream = torch.cuda.Stream()
    pv = api.Povu(device=local, stream=stream.cuda_stream)
    s = api.Session.from_graph(pv, g)
    del g
    n_comp = s.shard(rank, world) if world > 1 else None
    v, e, h, st = s.sizes()
    # A region is called only if EVERY selected reference path crosses it (color.cpp:104-163), so with one reference path per
    # chromosome the reference has to be run once per chromosome (-P 'S001#1#chr<c>'); so is this: one decompose + index per
    # step, then one call per owned component against that component's reference path.
    ga = s.graph_arrays(labels=False)
    own = [n for k, n in enumerate(ga["path_names"]) if n.startswith("S001#1#") and ga["step_off"][k + 1] > ga["step_off"][k]]
    del ga
    nrec = 0

    walls = []

    def step():
        n = 0
        t0 = time.perf_counter()
        s.hot_path([own[0]], keep_resident=True)
        n += int(s.lib.pvgpu_session_record_count(s.h))
        t1 = time.perf_counter()
        for pfx in own[1:]:
            s.call([pfx])
            n += int(s.lib.pvgpu_session_record_count(s.h))
        walls.append(((t1 - t0) * 1e3, (time.perf_counter() - t1) * 1e3))
        return n

    with torch.cuda.stream(stream):
        for _ in range(a.warmup):
            step()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(a.steps):
            nrec = step()
        e1.record(stream)
        torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / a.steps
    nfl = int(s.lib.pvgpu_session_flubble_count(s.h))
    t = torch.tensor([ms], dtype=torch.float64, device="cuda")
    cnt = torch.tensor([nrec, nfl, e, st], dtype=torch.float64, device="cuda")
    share = torch.tensor([float(e + st + v)], dtype=torch.float64, device="cuda")
    shares = [torch.zeros_like(share) for _ in range(world)]
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(cnt, op=dist.ReduceOp.SUM)
        dist.all_gather(shares, share)
    else:
        shares = [share]
    if rank == 0:
        w = [float(x.item()) for x in shares]
        print(json.dumps({"metric": "decompose_call_hot_path_edges_per_sec", "workload": f"wg_25c: 25 components ~ GRCh38 chromosome sizes, {a.instances} template instances, {a.paths} paths",
                          "scaling": "strong", "n_gpus": world, "segments": V, "links": E, "path_steps": S, "components": n_comp or 25,
                          "ms_per_step_max_over_ranks": t.item(), "value": E / (t.item() * 1e-3), "unit": "edges/s",
                          "records_all_ranks": int(cnt[0].item()), "flubbles_all_ranks": int(cnt[1].item()), "links_all_ranks": int(cnt[2].item()),
                          "lpt_load_max_over_mean": max(w) / (sum(w) / len(w)), "rank0_host_wall_ms_hot_path_and_calls": [round(x, 2) for x in walls[-1]], "rank0_calls": len(own) - 1, "steps": a.steps, "warmup": a.warmup}))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
