"""Synthetic map geometries shared by the parity tests (seeded; small enough for the CPU oracle)."""
import io

import numpy as np

from pdb_eda_b200 import synthetic


def _noise(shape, seed, smooth=1.0):
    from scipy import ndimage
    rng = np.random.default_rng(seed)
    v = ndimage.gaussian_filter(rng.standard_normal(shape), smooth, mode="wrap")
    return (v / v.std()).astype(np.float32)


# name -> dict(values shape (ns, nr, nc), cell, intervals (x, y, z), crsStart (c, r, s), axisOrder)
GEOMETRIES = {
    # orthogonal, whole cell stored, identity axis order
    "ortho": dict(shape=(40, 36, 32), cell=(16.0, 18.0, 20.0, 90, 90, 90), intervals=(32, 36, 40), crsStart=(0, 0, 0),
                  axisOrder=(1, 2, 3)),
    # orthogonal, permuted axes (col->Y, row->X), non-zero start, fewer stored voxels than intervals
    "perm": dict(shape=(30, 26, 28), cell=(24.0, 22.0, 26.0, 90, 90, 90), intervals=(48, 44, 52), crsStart=(-5, 3, 7),
                 axisOrder=(2, 1, 3)),
    # orthogonal, more stored voxels than intervals on two axes (the map repeats), section axis carries x
    "over": dict(shape=(36, 40, 24), cell=(15.0, 16.0, 12.0, 90, 90, 90), intervals=(30, 32, 24), crsStart=(2, -3, 0),
                 axisOrder=(2, 3, 1)),
    # hexagonal cell (gamma = 120): skewed path, BLAS-ordered mat-vec
    "hex": dict(shape=(40, 30, 30), cell=(15.0, 15.0, 30.0, 90, 90, 120), intervals=(30, 30, 40), crsStart=(0, 0, 0),
                axisOrder=(1, 2, 3)),
    # triclinic, permuted, shifted
    "tric": dict(shape=(28, 32, 30), cell=(17.0, 19.0, 21.0, 81.0, 97.0, 108.0), intervals=(34, 38, 42), crsStart=(4, -6, 3),
                 axisOrder=(3, 1, 2)),
}


def make_case(name, seed=11):
    g = GEOMETRIES[name]
    values = _noise(g["shape"], seed)
    data = synthetic.ccp4Bytes(values, g["cell"], g["intervals"], crsStart=g["crsStart"], axisOrder=g["axisOrder"])
    return data, values


def parse_with(ccp4_module, data, pdbid="case"):
    return ccp4_module.parse(io.BytesIO(data), pdbid)


def random_atoms(dm, n, seed, margin=2.0):
    """n float32 xyz positions spread over (and slightly beyond) the stored map, rounded to 3 decimals."""
    rng = np.random.default_rng(seed)
    h = dm.header
    crs = np.stack([rng.uniform(-margin, h.ncrs[a] + margin, n) for a in range(3)], axis=1)
    xyz = np.array([np.asarray(h.crs2xyzCoord([float(c) for c in row]), dtype=np.float64) for row in crs])
    return np.round(xyz, 3).astype(np.float32)


def random_geometry(seed):
    """A random cell (orthogonal / monoclinic / triclinic by seed % 3), axis order, start and stored extent (fewer or more
    voxels than intervals).  Returns (ccp4 bytes, values)."""
    import itertools
    rng = np.random.default_rng(1000 + seed)
    intervals = tuple(int(v) for v in rng.integers(20, 44, 3))
    cell_len = [iv * float(rng.uniform(0.35, 0.8)) for iv in intervals]
    if seed % 3 == 0:
        angles = (90.0, 90.0, 90.0)
    elif seed % 3 == 1:
        angles = (90.0, float(rng.uniform(95, 120)), 90.0)
    else:
        angles = tuple(float(v) for v in rng.uniform(70, 115, 3))
    order = list(itertools.permutations((1, 2, 3)))[int(rng.integers(6))]
    axes = [a - 1 for a in order]
    ncrs = [int(intervals[axes[k]] * rng.uniform(0.6, 1.3)) for k in range(3)]
    start = tuple(int(v) for v in rng.integers(-9, 10, 3))
    values = _noise((ncrs[2], ncrs[1], ncrs[0]), 2000 + seed)
    return synthetic.ccp4Bytes(values, tuple(cell_len) + angles, intervals, crsStart=start, axisOrder=order), values
