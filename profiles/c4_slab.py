#!/usr/bin/env python3
"""BASELINE.json config 4: one 1024^3 Fo-Fc map, slab-partitioned across the ranks of a torchrun launch, halo merge over
NCCL; timed on the device (max over ranks) and checked on rank 0 against the whole-map labelling.
usage: torchrun --nproc-per-node N profiles/c4_slab.py [n=1024]"""
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pdb_eda_b200 import _device, ccp4, slab, synthetic  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
vol = synthetic.smoothNoiseMapDevice(n, seed=4, device=dev)           # every rank generates the same map, keeps its slab
hdr = ccp4.DensityHeader.fromFileHeader(synthetic.ccp4Header((n, n, n), (n * 0.5,) * 3 + (90, 90, 90), (n, n, n)))
s0, s1 = slab.slabRanges(n, world)[rank]
mine = vol[s0:s1].contiguous()
cut = 3.0
for _ in range(2):
    parts = slab.labelSlabDistributed(hdr, mine, s0, s1, cut, -cut)
dist.barrier()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
reps = 3
for _ in range(reps):
    parts = slab.labelSlabDistributed(hdr, mine, s0, s1, cut, -cut)
e1.record()
torch.cuda.synchronize()
t = torch.tensor([e0.elapsed_time(e1) / reps], dtype=torch.float64, device=dev)
dist.all_reduce(t, op=dist.ReduceOp.MAX)
ok = True
if rank == 0:
    whole_dev = _device.DeviceMap(_device.geom_from_header(hdr), vol.reshape(-1))
    whole = whole_dev.blob_label(cut, -cut)
    e0.record()
    whole = whole_dev.blob_label(cut, -cut)
    e1.record()
    torch.cuda.synchronize()
    for w, p in zip(whole, parts):
        ok = ok and w["n_blobs"] == p["n_blobs"] and torch.allclose(w["stats"], p["stats"], rtol=1e-9, atol=1e-9)
    print("C4 n=%d world=%d: slab labelling %.2f ms (%.3g blob-CCL voxels/s), whole map on one GPU %.2f ms; blobs %d green / %d red; "
          "matches whole-map labelling: %s" % (n, world, t.item(), n ** 3 / (t.item() * 1e-3), e0.elapsed_time(e1), parts[0]["n_blobs"],
                                                 parts[1]["n_blobs"], ok))
dist.barrier()
dist.destroy_process_group()
