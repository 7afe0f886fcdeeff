#!/usr/bin/env python3
"""BASELINE.json config 3 on one GPU: a pool of synthetic structures (mixed sizes / space groups) analysed in batches
(multi.PoolShard).  usage: python profiles/c3_pool.py [structures] [passes] [maxAtomsPerBatch]"""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pdb_eda_b200 import _device, _lib, ccp4, cloudBatch, multi, synthetic  # noqa: E402

if os.environ.get("PE_LIB"):  # A/B runs of two builds of the library (tuning only)
    _lib.LIB_PATH = os.path.abspath(os.environ["PE_LIB"])

S = int(sys.argv[1]) if len(sys.argv) > 1 else 256
passes = int(sys.argv[2]) if len(sys.argv) > 2 else 3
maxAtoms = int(sys.argv[3]) if len(sys.argv) > 3 else 1 << 21
params = synthetic.defaultParams()
electronsOf = np.array([synthetic.ALA_ELECTRONS["ALA_" + a] for a, _, _ in synthetic._ALA_ATOMS])

t0 = time.perf_counter()
entries = []
for sp in synthetic.poolSpec(S):
    n, cell = sp["n"], sp["cell"]
    coords, bf = synthetic.fastPolyAla(sp["residues"], cell, sp["seed"])
    table = synthetic.polyAlaTable(coords, bf, params)
    rho = synthetic.densityMapDevice(coords, np.tile(electronsOf, sp["residues"]), n, cell, sp["seed"] + 1)
    hdr = ccp4.DensityHeader.fromFileHeader(synthetic.ccp4Header((n, n, n), cell, (n, n, n)))
    dmap = _device.DeviceMap(_device.geom_from_header(hdr), rho.reshape(-1))
    m, s = dmap.mean_std()
    entries.append((sp["index"], cloudBatch.MapRef(dmap, m + 1.5 * s, hdr.unitVolume, "s%05d" % sp["index"]), table))
torch.cuda.synchronize()
print("pool of %d structures built in %.1f s; %.2f GB of maps, %d atoms" % (S, time.perf_counter() - t0, sum(e[1].deviceMap.rho.numel() for e in entries) * 4 / 1e9,
                                                                   sum(len(e[2]) for e in entries)))
shard = multi.PoolShard(entries, params, maxAtoms)
print("batches:", len(shard.batches), "unsupported:", sum(not e[2].supported for e in entries))
out = shard.analyze()
torch.cuda.synchronize()
_lib.profile(True, reset=True)
t0 = time.perf_counter()
for _ in range(passes):
    out = shard.analyze()
torch.cuda.synchronize()
dt = (time.perf_counter() - t0) / passes
prof = _lib.profile()
_lib.profile(False)
print("pass: %.1f ms wall, %.1f structures/s, entries %d" % (dt * 1e3, S / dt, sum(b.nEntries for b in shard.batches)))
tot = 0.0
for name, (count, ms) in sorted(prof.items(), key=lambda kv: -kv[1][1]):
    print("  %-24s %5d launches %10.1f us/pass" % (name, count, ms * 1e3 / passes))
    tot += ms
print("  kernels total %.1f ms/pass" % (tot / passes))
t0 = time.perf_counter()
for _ in range(passes):
    shard.launch()
torch.cuda.synchronize()
t1 = time.perf_counter()
for _ in range(passes):
    shard.pack()
t2 = time.perf_counter()
print("launch %.1f ms, pack %.1f ms per pass" % ((t1 - t0) / passes * 1e3, (t2 - t1) / passes * 1e3))
c = out["cumulative"]
print("structures ok %d, ratio %.6f, voxels %d" % (c["structures"], c["density_electron_ratio"], c["num_voxels_aggregated"]))
print("medianDiffs", {k[:6]: round(v, 5) for k, v in out["medianDiffs"].items()})
if os.environ.get("CHECK_HASH"):
    # the pair path against the hash-table path over the whole pool: every per-structure number and per-atom flag
    import ctypes
    def arrays():
        outs = []
        for b in shard.batches:
            b.launch()
            arr = b.collectArrays()
            outs.append((b.atomRows().copy(), {k: np.asarray(v).copy() for k, v in arr.items() if k not in ("medians", "unitVolume")}))
        return outs
    a0 = arrays()
    os.environ["PE_CLOUD_FORCE_HASH"] = "1"
    a1 = arrays()
    os.environ["PE_CLOUD_FORCE_HASH"] = "0"
    same = all(np.array_equal(x[0], y[0], equal_nan=True) and all(np.array_equal(x[1][k], y[1][k], equal_nan=True) for k in x[1])
               for x, y in zip(a0, a1))
    print("pair path == hash-table path over the pool (per-atom records and per-structure totals, bit for bit):", same)
