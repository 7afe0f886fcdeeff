#!/usr/bin/env python3
"""BASELINE.json config 3 (scaled down): synthetic structures of mixed sizes / space groups sharded over the ranks of a
torchrun launch (and, per rank, over a pool of host worker processes); cumulative statistics all-reduced.
usage: torchrun --nproc-per-node N profiles/c3_multi.py [n_structures=64] [workers_per_rank=1]"""
import io
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pdb_eda_b200 import densityAnalysis, multi, structure, synthetic  # noqa: E402


def build(spec):
    n, sg, i = spec
    cell = (n * 0.5,) * 3 + (90.0, 90.0, 90.0)
    st = synthetic.polyAlaStructure(max(30, n ** 3 // 1750), (0, 0, 0), cell[:3], seed=100 + i, residuesPerChain=200)
    a, b = synthetic.mapPair(st, (n, n, n), cell, seed=200 + i)
    text = structure.formatPDB(st, remark290=synthetic.cartesianOperators(sg, cell), cell=cell, spaceGroup=sg)
    return text, synthetic.ccp4Bytes(a, cell, (n, n, n)), synthetic.ccp4Bytes(b, cell, (n, n, n))


def loader(item):
    an = densityAnalysis.fromFile(io.StringIO(item[0]), io.BytesIO(item[1]), io.BytesIO(item[2]))
    if an:
        _ = an.greenBlobList, an.redBlobList                           # the blob lists are part of the per-structure work
    return an


def main():
    count = int(sys.argv[1]) if len(sys.argv) > 1 else 64
    workers = int(sys.argv[2]) if len(sys.argv) > 2 else 1
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    densityAnalysis.setGlobals(synthetic.defaultParams())
    rng = np.random.default_rng(3)
    sizes = rng.choice([64, 96, 128], count)
    groups = rng.choice(["P 1", "P 1 21 1", "P 21 21 21", "P 43 21 2"], count)
    specs = [(int(n), str(g), i) for i, (n, g) in enumerate(zip(sizes, groups))]
    costs = [n ** 3 for n, _, _ in specs]
    mine = set(multi.shardStructures(costs, world)[rank])
    items = [build(s) if s[2] in mine else None for s in specs]        # every rank synthesises only its own share
    multi.runMultipleStructures(items[:0], loader, [], None, dev)       # warm-up of the collectives
    loader(next(it for it in items if it is not None))                  # and of this process's CUDA context
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    summary = multi.runMultipleStructures(items, loader, costs, None, dev, workers=workers)
    torch.cuda.synchronize()
    dt = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(dt, op=dist.ReduceOp.MAX)
    if rank == 0:
        c = summary["cumulative"]
        print("C3 (scaled): %d structures (64^3-128^3, mixed space groups), %d GPU(s) x %d host worker(s), %d host cores: %.2f s -> "
              "%.1f structures/s; analysed %d, pooled ratio %.6g, voxels aggregated %d"
              % (count, world, workers, os.cpu_count(), dt.item(), count / dt.item(), c["structures"], c["density_electron_ratio"],
                 c["num_voxels_aggregated"]))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
