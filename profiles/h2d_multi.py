#!/usr/bin/env python3
"""Host -> device copy ceiling of the box with N ranks copying at once (what bounds bench.py's `e2e` at N > 1): every rank
binds to the cores NVML reports for its GPU (bench.bind_to_gpu), allocates its pinned buffer after binding and copies it to its
GPU in a loop; all ranks start together.  Prints per-rank and aggregate GB/s and the CPU / NUMA layout the ranks see.
usage: python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 --master-port P profiles/h2d_multi.py [MiB=1024]"""
import os
import subprocess
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402

rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
mib = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
aff = bench.bind_to_gpu(local)
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
host = torch.empty(mib << 20, dtype=torch.uint8).pin_memory()
host.fill_(1)
dev = torch.empty_like(host, device="cuda")
for _ in range(2):
    dev.copy_(host, non_blocking=True)
torch.cuda.synchronize()
reps = 10
out = {}
for mode in ("alone", "together"):
    if mode == "alone":                      # one rank at a time
        ms = None
        for r in range(world):
            if world > 1:
                dist.barrier()
            if r == rank:
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                for _ in range(reps):
                    dev.copy_(host, non_blocking=True)
                e1.record()
                torch.cuda.synchronize()
                ms = e0.elapsed_time(e1) / reps
    else:
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            dev.copy_(host, non_blocking=True)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / reps
    gbs = torch.tensor([host.numel() / ms / 1e6], dtype=torch.float64, device="cuda")
    if world > 1:
        allg = [torch.zeros_like(gbs) for _ in range(world)]
        dist.all_gather(allg, gbs)
        out[mode] = [round(float(x.item()), 1) for x in allg]
    else:
        out[mode] = [round(float(gbs.item()), 1)]
if rank == 0:
    print("ranks %d, %d MiB per copy, pinned buffers allocated after binding; affinity of rank 0: %s" % (world, mib, aff))
    print("GB/s per rank, one rank at a time : %s" % out["alone"])
    print("GB/s per rank, all ranks together : %s  (aggregate %.1f)" % (out["together"], sum(out["together"])))
    for cmd in (["nvidia-smi", "topo", "-m"], ["lscpu"]):
        try:
            txt = subprocess.run(cmd, capture_output=True, text=True, timeout=20).stdout
            keep = [l for l in txt.splitlines() if cmd[0] == "nvidia-smi" or any(k in l for k in ("NUMA", "Socket", "Model name", "CPU(s):"))]
            print("\n".join(keep[:40]))
        except Exception as exc:
            print(cmd, "failed:", exc)
if world > 1:
    dist.destroy_process_group()
