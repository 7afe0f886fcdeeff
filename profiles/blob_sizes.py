#!/usr/bin/env python3
"""pe_blob_label (threshold + sparse labelling, both signs at +-3 sigma) on smoothed-noise maps of several sizes:
per-kernel device times from the library's event profiler and the phase split of the sparse stage.
usage: python profiles/blob_sizes.py [n ...]   (default 384 768 1024)"""
import ctypes
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pdb_eda_b200 import _device, _lib, ccp4, synthetic  # noqa: E402
from pdb_eda_b200._device import _ptr, _stream  # noqa: E402

if os.environ.get("PE_LIB"):
    _lib.LIB_PATH = os.path.abspath(os.environ["PE_LIB"])
lib = _lib.load()
for n in [int(a) for a in sys.argv[1:]] or [384, 768, 1024]:
    vol = synthetic.smoothNoiseMapDevice(n, seed=4)
    hdr = ccp4.DensityHeader.fromFileHeader(synthetic.ccp4Header((n, n, n), (n * 0.5,) * 3 + (90, 90, 90), (n, n, n)))
    dev = _device.DeviceMap(_device.geom_from_header(hdr), vol.reshape(-1))
    m, s = dev.mean_std()
    cut = float(torch.tensor(m + 3 * s, dtype=torch.float32))
    g = dev.geom
    cap_v, cap_b = n ** 3 // 64, n ** 3 // 256
    counts = torch.zeros(5, dtype=torch.int64, device="cuda")
    key = torch.empty(2 * cap_v, dtype=torch.int32, device="cuda")
    value = torch.empty(2 * cap_v, dtype=torch.float32, device="cuda")
    label = torch.empty(2 * cap_v, dtype=torch.int32, device="cuda")
    stats = torch.empty((2 * cap_b, 8), dtype=torch.float64, device="cuda")
    ws = torch.empty(int(lib.pe_blob_workspace_bytes(ctypes.byref(g), cap_v)), dtype=torch.uint8, device="cuda")

    def run():
        _lib.check(lib.pe_blob_label(ctypes.byref(g), _ptr(dev.rho), ctypes.c_float(cut), ctypes.c_float(-cut), cap_v, cap_b,
                                     _ptr(counts), _ptr(key), _ptr(value), _ptr(label), _ptr(stats), _ptr(ws), _stream()), "pe_blob_label")

    for _ in range(3):
        run()
    torch.cuda.synchronize()
    _lib.profile(True, reset=True)
    reps = 5
    for _ in range(reps):
        run()
    torch.cuda.synchronize()
    prof = _lib.profile()
    _lib.profile(False)
    t = (ctypes.c_ulonglong * 12)()
    lib.pe_blob_stage_times(t)
    t = list(t)
    c = counts.tolist()
    tot_us = sum(ms for _, ms in prof.values()) * 1e3 / reps
    print("n=%d: %d + %d foreground voxels, %d + %d blobs, overflow %d" % (n, c[0], c[2], c[1], c[3], c[4]))
    for name, (cnt, ms) in sorted(prof.items(), key=lambda kv: -kv[1][1]):
        print("   %-26s %9.1f us/launch" % (name, ms * 1e3 / cnt))
    print("   sparse phases P1..P4 (us): %s" % [round((t[i + 1] - t[i]) / 1e3, 1) for i in range(4)])
    print("   blob-CCL path: %.1f us, %.3g voxels/s, %.0f GB/s algorithmic (4 N_vox + 8 N_fg + 64 N_blob) = %.1f %% of 6560 GB/s"
          % (tot_us, n ** 3 / (tot_us * 1e-6), (4.0 * n ** 3 + 8.0 * (c[0] + c[2]) + 64.0 * (c[1] + c[3])) / (tot_us * 1e-6) / 1e9,
             (4.0 * n ** 3 + 8.0 * (c[0] + c[2]) + 64.0 * (c[1] + c[3])) / (tot_us * 1e-6) / 1e9 / 65.60))
    del vol, dev, key, value, label, stats, ws
    torch.cuda.empty_cache()
