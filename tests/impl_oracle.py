"""Adapter: the CPU restatement oracle (oracle/pe_oracle.c) behind the interface golden_checks.py drives."""
import numpy as np

from oracle import orc
from pdb_eda_b200 import _blas


class OracleImpl:
    def __init__(self, dm):
        self.header = dm.header
        self.rho = np.ascontiguousarray(dm.densityArray, dtype=np.float32)
        self.g = orc.geom(dm.header, dm.origin, mv=_blas.probe())
        self.orthogonal = bool(self.g.orthogonal)

    def xyz2crs(self, xyz):
        return orc.xyz2crs(self.g, xyz)

    def crs2xyz(self, crs):
        return orc.crs2xyz(self.g, crs)

    def point_density(self, crs):
        return orc.point_density(self.g, self.rho, crs)

    def mean_std(self):
        v = self.rho.astype(np.float64)
        return float(np.mean(v)), float(np.std(v))

    def sum_abs(self, cut):
        return orc.sum_abs(self.rho, cut)

    def sphere_lists(self, atoms, radii, cutoff):
        lists = [orc.sphere_list(self.g, self.rho, a, r, cutoff) for a, r in zip(np.asarray(atoms, dtype=np.float64), radii)]
        off = np.zeros(len(lists) + 1, dtype=np.int64)
        off[1:] = np.cumsum([len(l) for l in lists])
        return (np.concatenate(lists) if lists else np.zeros((0, 3), np.int32)), off

    def sphere_sums(self, atoms, radii, group_start, cp, cn):
        atoms = np.asarray(atoms, dtype=np.float64)
        radii = np.asarray(radii, dtype=np.float32)
        if group_start is None:
            group_start = np.arange(len(atoms) + 1)
        rows = []
        for k in range(len(group_start) - 1):
            s, e = group_start[k], group_start[k + 1]
            rows.append(orc.sphere_union_sums(self.g, self.rho, atoms[s:e], radii[s:e], cp, cn))
        return np.array(rows).reshape(-1, 7)

    def sphere_clouds(self, atoms, radii, cutoff):
        count, sizes, totals = [], [], []
        for a, r in zip(np.asarray(atoms, dtype=np.float64), radii):
            crs = orc.sphere_list(self.g, self.rho, a, r, cutoff)
            label, n = orc.cluster_crs(crs)
            st = orc.blob_stats(self.g, self.rho, crs, label, n)
            count.append(n)
            sizes.extend(st[:, 0].astype(np.int64).tolist())
            totals.extend(st[:, 1].tolist())
        return np.array(count), np.array(sizes, dtype=np.int64), np.array(totals)

    def full_blobs(self, cp, cn):
        out = []
        for c in (cp, cn):
            crs, label, n = orc.full_blobs(self.g, self.rho, c)
            out.append((crs, label, orc.blob_stats(self.g, self.rho, crs, label, n)))
        return out

    def cluster(self, crs):
        return orc.cluster_crs(crs)[0]

    def symmetry(self, xyz, ops, shift, lo, hi):
        return orc.symmetry(self.g, xyz, ops, shift, lo, hi)

    def nearest(self, cents, coords):
        return orc.nearest(cents, coords)
