#!/usr/bin/env python3
"""Times threshold_bitmap_kernel alone (L2 flushed between launches) for one PE_BMP_CFG; used for tuning."""
import os, sys, ctypes
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pdb_eda_b200 import _device, _lib, ccp4, synthetic
n = int(sys.argv[1]) if len(sys.argv) > 1 else 384
hdr = ccp4.DensityHeader.fromFileHeader(synthetic.ccp4Header((n, n, n), (n * 0.5,) * 3 + (90, 90, 90), (n, n, n)))
dev = _device.DeviceMap(_device.geom_from_header(hdr), synthetic.smoothNoiseMapDevice(n, seed=2).reshape(-1))
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
dev.blob_label(3.0, -3.0)
_lib.profile(True, reset=True)
for _ in range(10):
    flush.zero_()
    dev.blob_label(3.0, -3.0)
torch.cuda.synchronize()
prof = _lib.profile()
c, ms = prof["threshold_bitmap_kernel"]
us = ms * 1e3 / c
print("cfg=%s n=%d threshold_bitmap_kernel %.2f us/launch -> %.0f GB/s (%.1f%% of 6560)" % (os.environ.get("PE_BMP_CFG", "0"), n, us, 4.0 * n ** 3 / us / 1e3, 4.0 * n ** 3 / us / 1e3 / 65.6))
