"""Adapter: the CUDA library through its C ABI (pdb_eda_b200._device / cutils) behind the interface
golden_checks.py drives.  Everything here runs on the GPU; nothing falls back to the oracle."""
import numpy as np
import torch

from pdb_eda_b200 import _device, cutils


class CudaImpl:
    def __init__(self, dm):
        self.dm = dm
        self.header = dm.header
        self.dev = dm.deviceMap
        self.orthogonal = bool(self.dev.geom.orthogonal)

    def xyz2crs(self, xyz):
        return self.dev.xyz2crs(np.asarray(xyz, dtype=np.float64)).cpu().numpy()

    def crs2xyz(self, crs):
        return self.dev.crs2xyz(crs).cpu().numpy()

    def point_density(self, crs):
        val, ok = self.dev.point_density(crs)
        return val.cpu().numpy().astype(np.float64), ok.cpu().numpy().astype(bool)

    def mean_std(self):
        return self.dev.mean_std()

    def sum_abs(self, cut):
        return self.dev.sum_abs(cut)

    def sphere_lists(self, atoms, radii, cutoff):
        res = self.dev.sphere_lists(np.asarray(atoms, dtype=np.float64), np.asarray(radii, dtype=np.float32), cutoff)
        return res["crs"].cpu().numpy(), res["offset"].cpu().numpy()

    def sphere_sums(self, atoms, radii, group_start, cp, cn):
        out = self.dev.sphere_sums(np.asarray(atoms, dtype=np.float64), np.asarray(radii, dtype=np.float32), group_start, cp, cn)
        return out.cpu().numpy()

    def sphere_clouds(self, atoms, radii, cutoff):
        res = self.dev.sphere_lists(np.asarray(atoms, dtype=np.float64), np.asarray(radii, dtype=np.float32), cutoff,
                                    want_values=True, want_labels=True)
        n = len(atoms)
        atom = res["atom"].long()
        label = res["label"].long()
        count = torch.zeros(n, dtype=torch.int64, device=atom.device)
        if len(atom):
            count.scatter_reduce_(0, atom, label + 1, reduce="amax")
        cloud_off = torch.zeros(n + 1, dtype=torch.int64, device=atom.device)
        cloud_off[1:] = torch.cumsum(count, 0)
        gid = (cloud_off[atom] + label).to(torch.int32)
        ncl = int(cloud_off[-1].item())
        stats = cutils.crsStats(self.dm, res["crs"], gid, None, ncl).cpu().numpy()
        return count.cpu().numpy(), stats[:, 0].astype(np.int64), stats[:, 1]

    def full_blobs(self, cp, cn):
        out = []
        for part in self.dev.blob_label(cp, cn):
            out.append((part["crs"].cpu().numpy(), part["label"].cpu().numpy(), part["stats"].cpu().numpy()))
        return out

    def cluster(self, crs):
        return _device.cluster_crs(crs)[0].cpu().numpy()

    def symmetry(self, xyz, ops, shift, lo, hi):
        a, i, x = self.dev.symmetry_expand(xyz, ops, shift, lo, hi)
        return a.cpu().numpy(), i.cpu().numpy(), x.cpu().numpy()

    def nearest(self, cents, coords):
        i, d = _device.nearest_atom(cents, coords)
        return i.cpu().numpy(), d.cpu().numpy()
