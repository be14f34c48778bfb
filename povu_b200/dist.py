"""Multi-GPU gfa2vcf: one process per GPU, weakly connected components sharded across ranks
(SURVEY.md 8e).  torch.distributed is plumbing only: the ranks exchange nothing while computing; at the end
the per-component region counts are all-gathered (to number the regions globally, so that the chunked record
order equals the single-GPU / reference order, genomics.cpp:114-156) and the record lines are gathered to rank 0.
Works with backend "nccl" (GPU box) and "gloo" (CPU tests with the host rehearsal library)."""
from __future__ import annotations

import os

from . import api


def gfa2vcf(pv, gfa_path, prefixes, chunk_size=100, with_date=False, pvst_dir=None):
    """Returns the combined VCF text on rank 0 (None elsewhere).  If pvst_dir is given every rank writes the
    <component_id>.pvst files of the components it owns there."""
    import torch.distributed as dist
    rank, world = (dist.get_rank(), dist.get_world_size()) if dist.is_initialized() else (0, 1)
    s = api.Session(pv, gfa_path)
    header = s_header = None
    if rank == 0:
        # the header needs only the path table (kept whole on every rank)
        pass
    if world > 1:
        s.shard(rank, world)
    s.decompose()
    if pvst_dir is not None:
        os.makedirs(pvst_dir, exist_ok=True)
        s.write_pvst(pvst_dir)
    s.call(prefixes)
    st = s.call_stats()
    lines, keys = s.record_lines()
    regions = s.tree_regions()
    if world == 1:
        return s.vcf_text(chunk_size=chunk_size, with_date=with_date), st
    all_regions = [None] * world
    dist.all_gather_object(all_regions, regions)
    merged = {}
    for d in all_regions:
        merged.update(d)
    base, tot = {}, 0
    for comp in sorted(merged):          # forest files are visited in ascending component id
        base[comp] = tot
        tot += merged[comp]
    last_chunk = max(tot - 1, 0) // chunk_size
    payload = []
    for i, (c, o, r) in enumerate(keys.tolist()):
        if c == 0xFFFFFFFF:      # SUBR records: emitted with the final chunk, after its ordinary records of that ref
            payload.append((last_chunk, int(r), tot, i, lines[i]))
        else:
            payload.append(((base[int(c)] + int(o)) // chunk_size, int(r), base[int(c)] + int(o), i, lines[i]))
    gathered = [None] * world if rank == 0 else None
    dist.gather_object((payload, st), gathered, dst=0)
    if rank != 0:
        return None, st
    recs = []
    tot_st = {k: 0 for k in st}
    for p, t in gathered:
        recs += p
        for k in t:
            tot_st[k] = max(tot_st[k], t[k]) if k == "max_region" else tot_st[k] + t[k]
    recs.sort(key=lambda x: x[:4])
    head = s.vcf_text(chunk_size=chunk_size, with_date=with_date)
    head = "".join(l for l in head.splitlines(True) if l.startswith("#"))
    return head + "".join(r[4] for r in recs), tot_st
