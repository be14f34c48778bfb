"""Worker of the world_size-2 tests: `python dist_worker.py <lib or 'cuda'> <gfa> <prefix> <chunk> <out_dir>`
under torch.distributed.run.  Rank 0 writes <out_dir>/out.vcf and stats.json; every rank writes its .pvst files."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    lib, gfa, prefix, chunk, out = sys.argv[1:6]
    import torch.distributed as dist
    from povu_b200 import api, dist as pvdist
    rank = int(os.environ["RANK"])
    if lib == "cuda":
        import torch
        torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", "0")))
        dist.init_process_group("nccl")
        pv = api.Povu(device=int(os.environ.get("LOCAL_RANK", "0")))
    else:
        dist.init_process_group("gloo")
        pv = api.Povu(lib_path=lib)
    text, st = pvdist.gfa2vcf(pv, gfa, [prefix], chunk_size=int(chunk), pvst_dir=os.path.join(out, "forest"))
    if rank == 0:
        open(os.path.join(out, "out.vcf"), "w").write(text)
        json.dump(st, open(os.path.join(out, "stats.json"), "w"))
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
