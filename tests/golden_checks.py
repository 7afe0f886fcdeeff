"""Comparisons of an implementation (the CPU oracle, or the CUDA library through its C ABI) with the golden
vectors produced by the real reference (tests/golden/*.npz, made by tests/golden/make_golden.py).

Bars (BASELINE.json north_star): voxel sets, blob membership, labels, indices: bit-exact.  float64 sums and
statistics: 1e-9 relative.
"""
import io
import os

import numpy as np

from pdb_eda_b200 import ccp4 as my_ccp4
from pdb_eda_b200 import synthetic

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
CASES = ("ortho", "perm", "over", "hex", "tric")
RTOL = 1e-9


def load(name):
    z = np.load(os.path.join(GOLDEN_DIR, name + ".npz"))
    return {k: z[k] for k in z.files}


def header_and_bytes(gold):
    data = synthetic.ccp4Bytes(gold["values"], tuple(gold["cell"]), tuple(int(v) for v in gold["intervals"]),
                               crsStart=tuple(int(v) for v in gold["crsStart"]), axisOrder=tuple(int(v) for v in gold["axisOrder"]))
    dm = my_ccp4.parse(io.BytesIO(data), "golden")
    return dm, data


def close(a, b, rtol=RTOL, atol=0.0):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    assert a.shape == b.shape, (a.shape, b.shape)
    scale = np.maximum(np.abs(a), np.abs(b))
    bad = np.abs(a - b) > rtol * scale + atol
    assert not bad.any(), "max rel err %.3e at %s" % (float((np.abs(a - b) / np.maximum(scale, 1e-300)).max()), np.argwhere(bad)[:5].tolist())


def check_header(gold, header):
    """Host header geometry (pdb_eda/ccp4.py:158-286) is bit-identical to the reference's."""
    ints = np.array([header.ncrs, header.crsStart, header.xyzInterval, header.crsInterval, header.uniqueNcrs, header.map2xyz,
                     header.map2crs])
    assert np.array_equal(ints, gold["hdr_ints"])
    assert np.array_equal(np.asarray(header.origin, dtype=np.float64), gold["hdr_origin"])
    assert np.array_equal(np.asarray(header.gridLength, dtype=np.float64), gold["hdr_gridLength"])
    assert np.array_equal(np.asarray(header.orthoMat, dtype=np.float64), gold["hdr_ortho"])
    assert np.array_equal(np.asarray(header.deOrthoMat, dtype=np.float64), gold["hdr_deortho"])
    assert float(header.unitVolume) == float(gold["unitVolume"])


def check_conversions(gold, impl):
    assert np.array_equal(impl.xyz2crs(gold["atoms"].astype(np.float64)), gold["atoms_crs"])
    assert np.array_equal(impl.xyz2crs(gold["xyz_pts"]), gold["xyz_pts_crs"])
    xyz = impl.crs2xyz(gold["crs_pts"])
    if impl.orthogonal:
        assert np.array_equal(xyz, gold["crs_pts_xyz"])  # scalar arithmetic with a fully specified order
    else:
        close(xyz, gold["crs_pts_xyz"], rtol=1e-12, atol=1e-12)  # BLAS-ordered mat-vec


def check_points(gold, impl):
    val, valid = impl.point_density(gold["crs_pts"])
    assert np.array_equal(np.asarray(valid, dtype=bool), gold["crs_pts_valid"])
    assert np.array_equal(np.asarray(val, dtype=np.float64), gold["crs_pts_val"])


def check_mean_std(gold, impl):
    mean, std = impl.mean_std()
    close([std], [float(gold["std"])])
    assert abs(mean - float(gold["mean"])) <= 1e-9 * float(gold["std"])


def check_sum_abs(gold, impl):
    for cut, want in zip(gold["sum_abs_cut"], gold["sum_abs"]):
        close([impl.sum_abs(float(cut))], [want])


def check_sphere_lists(gold, impl):
    """getSphereCrsFromXyz per atom, three cutoffs: identical voxel lists in identical order."""
    cut = float(gold["sphere_cut"])
    for tag, c in (("zero", 0.0), ("pos", cut), ("neg", -cut)):
        crs, off = impl.sphere_lists(gold["atoms"], gold["radii"], c)
        assert np.array_equal(np.asarray(off, dtype=np.int64), gold["sphere_%s_off" % tag]), tag
        assert np.array_equal(crs, gold["sphere_%s_crs" % tag]), tag


def check_sphere_sums(gold, impl):
    cut = float(gold["sphere_cut"])
    rows = impl.sphere_sums(gold["atoms"], gold["radii"], None, cut, -cut)
    n_zero = np.diff(gold["sphere_zero_off"])
    n_pos = np.diff(gold["sphere_pos_off"])
    n_neg = np.diff(gold["sphere_neg_off"])
    assert np.array_equal(rows[:, 0].astype(np.int64), n_zero)
    assert np.array_equal(rows[:, 2].astype(np.int64), n_pos)
    assert np.array_equal(rows[:, 4].astype(np.int64), n_neg)
    close(rows[:, 1], gold["sphere_total_zero"], atol=1e-12)
    close(rows[:, 3], gold["sphere_total_pos"], atol=1e-12)
    assert np.array_equal(rows[:, 6] != 0, gold["sphere_valid"])


def check_sphere_unions(gold, impl):
    """Set-union semantics of getSphereCrsFromXyzList + the region density / discrepancy sums."""
    cut = float(gold["sphere_cut"])
    radii = np.full(len(gold["atoms"]), 2.0)
    rows = impl.sphere_sums(gold["atoms"], radii, gold["group_start"], cut, -cut)
    want = gold["union_rows"]
    for col in (0, 2, 4, 6):
        assert np.array_equal(rows[:, col], want[:, col]), col
    for col in (1, 3, 5):
        close(rows[:, col], want[:, col], atol=1e-11)


def check_clouds(gold, impl):
    """findAberrantBlobs(atom): per-atom clusters in createCrsLists order, sizes and total densities."""
    count, sizes, totals = impl.sphere_clouds(gold["atoms"], gold["radii"], float(gold["sphere_cut"]))
    assert np.array_equal(np.asarray(count, dtype=np.int64), gold["cloud_count"])
    assert np.array_equal(np.asarray(sizes, dtype=np.int64), gold["cloud_sizes"])
    close(totals, gold["cloud_totals"])


def check_blobs(gold, impl):
    """createFullBlobList(+-cutoff): voxel list order, blob membership and numbering bit-exact; aggregates 1e-9."""
    bcut = float(gold["blob_cut"])
    res = impl.full_blobs(bcut, -bcut)
    for tag, part in zip(("green", "red"), res):
        crs, label, stats = part
        assert np.array_equal(crs, gold[tag + "_crs"]), tag
        assert np.array_equal(label, gold[tag + "_label"]), tag
        n = stats[:, 0]
        close(stats[:, 1], gold[tag + "_total"])
        close(stats[:, 2:5] / stats[:, 1:2], gold[tag + "_centroid"], rtol=1e-9, atol=1e-9)
        close(stats[:, 5:8] / n[:, None], gold[tag + "_center"], rtol=1e-9, atol=1e-9)
        close(n * float(gold["unitVolume"]), gold[tag + "_volume"])


def check_cluster(gold, impl):
    label = impl.cluster(gold["arb_crs"])
    assert np.array_equal(label, gold["arb_label"])


def check_symmetry(gold, impl, dm):
    ops = gold["sym_ops"]
    box = gold["sym_box"]
    shift = np.array([np.dot(dm.header.orthoMat, (i, j, k)) for i in (-1, 0, 1) for j in (-1, 0, 1) for k in (-1, 0, 1)])
    lo = [box[0] - 5, box[2] - 5, box[4] - 5]
    hi = [box[1] + 5, box[3] + 5, box[5] + 5]
    atom, image, xyz = impl.symmetry(gold["atoms"].astype(np.float64), ops, shift, lo, hi)
    nops = len(ops)
    img, op = np.divmod(image, nops)
    symmetry = np.stack((img // 9 - 1, (img // 3) % 3 - 1, img % 3 - 1, op), axis=1)
    assert np.array_equal(atom, gold["sym_atom"])
    assert np.array_equal(symmetry, gold["sym_symmetry"])
    close(xyz, gold["sym_xyz"], rtol=1e-12, atol=1e-11)
    return xyz


def check_nearest(gold, impl):
    if len(gold["near_idx"]) == 0:
        return
    idx, dist = impl.nearest(gold["green_centroid"], gold["sym_xyz"])
    assert np.array_equal(idx, gold["near_idx"])
    close(dist, gold["near_dist"], rtol=1e-12)
