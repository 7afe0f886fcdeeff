#!/usr/bin/env python3
"""Generates tests/golden/<case>.npz by running the REAL reference (oracle/_ref: unmodified pdb_eda 2.7.1 with its
compiled Cython cutils) on the seeded synthetic maps of tests/cases.py.

Run where /root/reference (or a built oracle/_ref) exists:   python tests/golden/make_golden.py
The fixtures are committed; the GPU box has no reference source, so the gpu tests read these files.
"""
import io
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import cases  # noqa: E402
from oracle import build_ref, refload  # noqa: E402
from pdb_eda_b200 import synthetic  # noqa: E402

SPACE_GROUP = {"ortho": "P 21 21 21", "perm": "P 21 21 21", "over": "P 1 21 1", "hex": "P 65 2 2", "tric": "P 1"}


class _Atom:
    def __init__(self, coord):
        self.coord = coord


def ragged(lists):
    off = np.zeros(len(lists) + 1, dtype=np.int64)
    for i, l in enumerate(lists):
        off[i + 1] = off[i] + len(l)
    flat = np.array([c for l in lists for c in l], dtype=np.int32).reshape(-1, 3)
    return flat, off


def blob_arrays(blobs, crs_list):
    """Per-voxel blob number (in the reference's blob order) for the createFullCrsList order + per-blob aggregates."""
    index = {}
    for b, blob in enumerate(blobs):
        for crs in blob.crsList:
            index[tuple(crs)] = b
    label = np.array([index[tuple(c)] for c in crs_list], dtype=np.int32)
    total = np.array([b.totalDensity for b in blobs], dtype=np.float64)
    centroid = np.array([b.centroid for b in blobs], dtype=np.float64).reshape(-1, 3)
    center = np.array([b.coordCenter for b in blobs], dtype=np.float64).reshape(-1, 3)
    volume = np.array([b.volume for b in blobs], dtype=np.float64)
    return label, total, centroid, center, volume


def make(name, ref_ccp4, ref_cutils):
    g = cases.GEOMETRIES[name]
    data, values = cases.make_case(name)
    dm = cases.parse_with(ref_ccp4, data)
    h = dm.header
    rng = np.random.default_rng({"ortho": 101, "perm": 102, "over": 103, "hex": 104, "tric": 105}[name])
    out = dict(values=values, cell=np.array(g["cell"], dtype=np.float64), intervals=np.array(g["intervals"]),
               crsStart=np.array(g["crsStart"]), axisOrder=np.array(g["axisOrder"]))
    mean, std = float(dm.meanDensity), float(dm.stdDensity)
    out["mean"], out["std"] = mean, std
    out["unitVolume"] = float(h.unitVolume)
    out["hdr_origin"] = np.asarray(h.origin, dtype=np.float64)
    out["hdr_gridLength"] = np.asarray(h.gridLength, dtype=np.float64)
    out["hdr_ortho"] = np.asarray(h.orthoMat, dtype=np.float64)
    out["hdr_deortho"] = np.asarray(h.deOrthoMat, dtype=np.float64)
    out["hdr_ints"] = np.array([h.ncrs, h.crsStart, h.xyzInterval, h.crsInterval, h.uniqueNcrs, h.map2xyz, h.map2crs])

    # ---- conversions and point lookups
    atoms = cases.random_atoms(dm, 36, seed=rng.integers(1 << 30))
    out["atoms"] = atoms
    out["atoms_crs"] = np.array([h.xyz2crsCoord(a) for a in atoms], dtype=np.int32)
    pts64 = rng.uniform(-30, 60, (200, 3))
    out["xyz_pts"] = pts64
    out["xyz_pts_crs"] = np.array([h.xyz2crsCoord(list(p)) for p in pts64], dtype=np.int32)
    crs_pts = np.stack([rng.integers(-3 * h.crsInterval[a], 4 * h.crsInterval[a], 400) for a in range(3)], axis=1).astype(np.int32)
    out["crs_pts"] = crs_pts
    out["crs_pts_xyz"] = np.array([np.asarray(h.crs2xyzCoord([int(v) for v in c]), dtype=np.float64) for c in crs_pts])
    out["crs_pts_val"] = np.array([float(ref_cutils.getPointDensityFromCrs(dm, [int(v) for v in c])) for c in crs_pts])
    out["crs_pts_valid"] = np.array([bool(ref_cutils.testValidCrs(dm, [int(v) for v in c])) for c in crs_pts])

    # ---- spheres (per atom; three cutoffs); radii include float32-unfriendly values
    radii = rng.uniform(0.6, 1.3, len(atoms))
    radii[::7] = 2.1
    radii[3] = 0.75 - 1e-9
    out["radii"] = radii  # float64 as a caller would pass them; the seam narrows to float32
    cut = mean + 1.5 * std
    out["sphere_cut"] = cut
    for tag, c in (("zero", 0.0), ("pos", cut), ("neg", -cut)):
        lists = [ref_cutils.getSphereCrsFromXyz(dm, a, r, c) for a, r in zip(atoms, radii)]
        flat, off = ragged(lists)
        out["sphere_%s_crs" % tag], out["sphere_%s_off" % tag] = flat, off
    out["sphere_total_zero"] = np.array([dm.getTotalDensityFromXyz(a, r, 0) for a, r in zip(atoms, radii)])
    out["sphere_total_pos"] = np.array([dm.getTotalDensityFromXyz(a, r, cut) for a, r in zip(atoms, radii)])
    out["sphere_valid"] = np.array([bool(ref_cutils.testValidXyz(dm, a, r)) for a, r in zip(atoms, radii)])
    # per-atom clouds: findAberrantBlobs(atom) -> clusters in creation order
    cloud_sizes, cloud_totals = [], []
    for a, r in zip(atoms, radii):
        blobs = dm.findAberrantBlobs(a, r, cut)
        cloud_sizes.append([len(b.crsList) for b in blobs])
        cloud_totals.append([b.totalDensity for b in blobs])
    out["cloud_count"] = np.array([len(s) for s in cloud_sizes])
    out["cloud_sizes"] = np.array([x for s in cloud_sizes for x in s], dtype=np.int64)
    out["cloud_totals"] = np.array([x for s in cloud_totals for x in s], dtype=np.float64)

    # ---- unions of spheres (groups of 3 atoms, radius 2.0): the region density / discrepancy quantities
    gstart = np.arange(0, len(atoms) + 1, 3)
    out["group_start"] = gstart.astype(np.int32)
    rows = []
    for k in range(len(gstart) - 1):
        xyz = [atoms[i] for i in range(gstart[k], gstart[k + 1])]
        union = ref_cutils.getSphereCrsFromXyzList(dm, xyz, 2.0)
        green = dm.findAberrantBlobs(xyz, 2.0, cut)
        red = dm.findAberrantBlobs(xyz, 2.0, -cut)
        rows.append([len(union), sum(ref_cutils.getPointDensityFromCrs(dm, c) for c in union),
                     sum(len(b.crsList) for b in green), sum(b.totalDensity for b in green),
                     sum(len(b.crsList) for b in red), sum(b.totalDensity for b in red),
                     float(ref_cutils.testValidXyzList(dm, xyz, 2.0))])
    out["union_rows"] = np.array(rows, dtype=np.float64)

    # ---- whole-map blobs at +-(mean + 2.5 sigma)
    bcut = mean + 2.5 * std
    out["blob_cut"] = bcut
    for tag, c in (("green", bcut), ("red", -bcut)):
        crs_list = ref_cutils.createFullCrsList(dm, c)
        blobs = dm.createFullBlobList(c)
        label, total, centroid, center, volume = blob_arrays(blobs, crs_list)
        out[tag + "_crs"] = np.array(crs_list, dtype=np.int32).reshape(-1, 3)
        out[tag + "_label"] = label
        out[tag + "_total"], out[tag + "_centroid"], out[tag + "_center"], out[tag + "_volume"] = total, centroid, center, volume
    out["sum_abs"] = np.array([ref_cutils.sumOfAbs(dm.densityArray, c) for c in (bcut, 0.5, 1.5 + 1e-9)])
    out["sum_abs_cut"] = np.array([bcut, 0.5, 1.5 + 1e-9])

    # ---- createCrsLists on an arbitrary (shuffled, un-wrapped) voxel list
    union = sorted(ref_cutils.getSphereCrsFromXyzList(dm, [atoms[i] for i in range(0, 12)], 1.6, cut))
    perm = rng.permutation(len(union))
    arb = [union[i] for i in perm]
    clusters = ref_cutils.createCrsLists(arb) if arb else []
    index = {tuple(c): k for k, cl in enumerate(clusters) for c in cl}
    out["arb_crs"] = np.array(arb, dtype=np.int32).reshape(-1, 3)
    out["arb_label"] = np.array([index[tuple(c)] for c in arb], dtype=np.int32)

    # ---- symmetry atoms + nearest atom per blob
    ops = synthetic.cartesianOperators(SPACE_GROUP[name], g["cell"])
    ncrs = h.ncrs
    corners = [h.crs2xyzCoord([c, r, s]) for c in [0, ncrs[0] - 1] for r in [0, ncrs[1] - 1] for s in [0, ncrs[2] - 1]]
    xs, ys, zs = (sorted(float(p[k]) for p in corners) for k in range(3))
    sym = ref_cutils.createSymmetryAtoms([_Atom(a) for a in atoms], ops, h.orthoMat, xs, ys, zs)
    out["sym_ops"] = np.array(ops, dtype=np.float64)
    out["sym_box"] = np.array([xs[0], xs[-1], ys[0], ys[-1], zs[0], zs[-1]])
    sym_atom = []
    for s in sym:
        key = s.atom.coord.tobytes()
        sym_atom.append(next(i for i, a in enumerate(atoms) if a.tobytes() == key))
    out["sym_atom"] = np.array(sym_atom, dtype=np.int32)
    out["sym_symmetry"] = np.array([s.symmetry for s in sym], dtype=np.int32)
    out["sym_xyz"] = np.array([np.asarray(s.coord, dtype=np.float64) for s in sym])
    import scipy.spatial
    cents = out["green_centroid"]
    if len(cents):
        d = scipy.spatial.distance.cdist(cents, out["sym_xyz"])
        out["near_idx"] = np.argmin(d, axis=1).astype(np.int32)
        out["near_dist"] = d.min(axis=1)
    else:
        out["near_idx"] = np.zeros(0, np.int32)
        out["near_dist"] = np.zeros(0)
    return out


def main():
    build_ref.build()
    ref_ccp4, _, ref_cutils, _ = refload.load()
    for name in cases.GEOMETRIES:
        out = make(name, ref_ccp4, ref_cutils)
        path = os.path.join(HERE, name + ".npz")
        np.savez_compressed(path, **out)
        print(name, "->", path, os.path.getsize(path) // 1024, "KiB;", len(out["green_total"]), "green /",
              len(out["red_total"]), "red blobs;", len(out["sym_atom"]), "symmetry atoms")


if __name__ == "__main__":
    main()
