"""ctypes binding of the C restatement oracle (``oracle/pe_oracle.c``).  TEST INFRASTRUCTURE ONLY.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline / ``--impl reference`` legs may import
this module; the product (``pdb_eda_b200/``) never does.  Parity status of the oracle: PINNED against the real
reference (``oracle/_ref``) by ``tests/test_oracle_vs_reference.py`` and the fixtures under ``tests/golden``.
"""
import ctypes
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "_build", "libpe_oracle.so")


class OrcGeom(ctypes.Structure):
    """``struct orc_geom`` of pe_oracle.c (same layout as the product's pe_geom, declared independently)."""
    _fields_ = [("ncrs", ctypes.c_int32 * 3), ("crs_start", ctypes.c_int32 * 3), ("xyz_interval", ctypes.c_int32 * 3),
                ("crs_interval", ctypes.c_int32 * 3), ("unique_ncrs", ctypes.c_int32 * 3), ("map2xyz", ctypes.c_int32 * 3),
                ("map2crs", ctypes.c_int32 * 3), ("orthogonal", ctypes.c_int32), ("mv_perm", ctypes.c_int32 * 3),
                ("mv_fma", ctypes.c_int32), ("grid_length", ctypes.c_double * 3), ("origin", ctypes.c_double * 3),
                ("ortho", ctypes.c_double * 9), ("deortho", ctypes.c_double * 9)]


def build(force=False):
    if force or not os.path.isfile(LIB_PATH) or os.path.getmtime(LIB_PATH) < os.path.getmtime(os.path.join(HERE, "pe_oracle.c")):
        subprocess.run(["make", "-C", HERE, "-B"], check=True, stdout=subprocess.DEVNULL)
    return LIB_PATH


_lib = None
_P = ctypes.c_void_p
_G = ctypes.POINTER(OrcGeom)
_I32, _I64, _F32 = ctypes.c_int32, ctypes.c_int64, ctypes.c_float


def lib():
    global _lib
    if _lib is None:
        build()
        L = ctypes.CDLL(LIB_PATH)
        sig = {
            "orc_xyz2crs": (None, [_G, _P, _P]), "orc_crs2xyz": (None, [_G, _P, _P]),
            "orc_point_density_batch": (None, [_G, _P, _I64, _P, _P, _P]),
            "orc_sphere_box": (None, [_G, _P, _F32, _P, _P]),
            "orc_sphere_list": (_I64, [_G, _P, _P, _F32, _F32, _P, _I64]),
            "orc_sphere_union_sums": (None, [_G, _P, _I32, _P, _P, _F32, _F32, _P]),
            "orc_threshold_list": (_I64, [_G, _P, _F32, _P, _I64]),
            "orc_cluster_crs": (_I64, [_I64, _P, _P]),
            "orc_blob_stats": (None, [_G, _P, _I64, _P, _P, _I64, _P]),
            "orc_full_blobs": (ctypes.c_int, [_G, _P, _F32, _P, _P, _I64, _P]),
            "orc_symmetry": (_I64, [_G, _I32, _P, _I32, _P, _P, _P, _P, _P, _P, _P, _I64]),
            "orc_nearest": (None, [_I64, _P, _I64, _P, _P, _P]),
            "orc_sum_abs": (ctypes.c_double, [_P, _I64, _F32]),
            "orc_test_overlap": (ctypes.c_int, [_I64, _P, _I64, _P]),
            "orc_sphere_sums_batch": (None, [_G, _P, _I32, _P, _P, _F32, _P]),
        }
        for name, (res, args) in sig.items():
            fn = getattr(L, name)
            fn.restype = res
            fn.argtypes = args
        _lib = L
    return _lib


def geom(header, origin=None, mv=((1, 0, 2), 1)):
    """OrcGeom from a duck-typed DensityHeader (the reference's or the product's).  ``mv``: accumulation order of
    the host BLAS's 3x3 mat-vec, (perm, fma)."""
    g = OrcGeom()
    origin = header.origin if origin is None else origin
    for a in range(3):
        g.ncrs[a] = int(header.ncrs[a])
        g.crs_start[a] = int(header.crsStart[a])
        g.xyz_interval[a] = int(header.xyzInterval[a])
        g.crs_interval[a] = int(header.crsInterval[a])
        g.unique_ncrs[a] = int(header.uniqueNcrs[a])
        g.map2xyz[a] = int(header.map2xyz[a])
        g.map2crs[a] = int(header.map2crs[a])
        g.mv_perm[a] = mv[0][a]
        g.grid_length[a] = float(header.gridLength[a])
        g.origin[a] = float(origin[a])
    g.mv_fma = mv[1]
    g.orthogonal = 1 if (header.alpha == header.beta == header.gamma == 90) else 0
    o = np.asarray(header.orthoMat, dtype=np.float64).reshape(9)
    d = np.asarray(header.deOrthoMat, dtype=np.float64).reshape(9)
    for k in range(9):
        g.ortho[k] = float(o[k])
        g.deortho[k] = float(d[k])
    return g


def _p(a):
    return a.ctypes.data_as(ctypes.c_void_p) if a is not None else None


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def _f64(a, shape=None):
    a = np.ascontiguousarray(a, dtype=np.float64)
    return a.reshape(shape) if shape else a


def _i32(a, shape=None):
    a = np.ascontiguousarray(a, dtype=np.int32)
    return a.reshape(shape) if shape else a


def xyz2crs(g, xyz):
    xyz = _f64(xyz, (-1, 3))
    out = np.empty(xyz.shape, dtype=np.int32)
    for i in range(len(xyz)):
        lib().orc_xyz2crs(ctypes.byref(g), _p(xyz[i:i + 1]), _p(out[i:i + 1]))
    return out


def crs2xyz(g, crs):
    crs = _i32(crs, (-1, 3))
    out = np.empty(crs.shape, dtype=np.float64)
    for i in range(len(crs)):
        lib().orc_crs2xyz(ctypes.byref(g), _p(crs[i:i + 1]), _p(out[i:i + 1]))
    return out


def point_density(g, rho, crs):
    rho = _f32(rho).reshape(-1)
    crs = _i32(crs, (-1, 3))
    out = np.empty(len(crs), dtype=np.float64)
    valid = np.empty(len(crs), dtype=np.int32)
    lib().orc_point_density_batch(ctypes.byref(g), _p(rho), len(crs), _p(crs), _p(out), _p(valid))
    return out, valid


def sphere_box(g, xyz, radius):
    xyz = _f64(xyz, (3,))
    lo = np.empty(3, dtype=np.int32)
    dim = np.empty(3, dtype=np.int32)
    lib().orc_sphere_box(ctypes.byref(g), _p(xyz), float(np.float32(radius)), _p(lo), _p(dim))
    return lo, dim


def sphere_list(g, rho, xyz, radius, cutoff=0.0):
    rho = _f32(rho).reshape(-1)
    xyz = _f64(xyz, (3,))
    r, c = float(np.float32(radius)), float(np.float32(cutoff))
    n = lib().orc_sphere_list(ctypes.byref(g), _p(rho), _p(xyz), r, c, None, 0)
    out = np.empty((n, 3), dtype=np.int32)
    lib().orc_sphere_list(ctypes.byref(g), _p(rho), _p(xyz), r, c, _p(out), n)
    return out


def sphere_union_sums(g, rho, xyz, radius, cut_pos=0.0, cut_neg=0.0):
    rho = _f32(rho).reshape(-1)
    xyz = _f64(xyz, (-1, 3))
    radius = _f32(np.broadcast_to(np.asarray(radius, dtype=np.float32), (len(xyz),)))
    out = np.empty(7, dtype=np.float64)
    lib().orc_sphere_union_sums(ctypes.byref(g), _p(rho), len(xyz), _p(xyz), _p(radius), float(np.float32(cut_pos)),
                                float(np.float32(cut_neg)), _p(out))
    return out


def sphere_sums_batch(g, rho, xyz, radius, cutoff=0.0):
    rho = _f32(rho).reshape(-1)
    xyz = _f64(xyz, (-1, 3))
    radius = _f32(np.broadcast_to(np.asarray(radius, dtype=np.float32), (len(xyz),)))
    out = np.empty((len(xyz), 4), dtype=np.float64)
    lib().orc_sphere_sums_batch(ctypes.byref(g), _p(rho), len(xyz), _p(xyz), _p(radius), float(np.float32(cutoff)), _p(out))
    return out


def threshold_list(g, rho, cutoff):
    rho = _f32(rho).reshape(-1)
    c = float(np.float32(cutoff))
    n = lib().orc_threshold_list(ctypes.byref(g), _p(rho), c, None, 0)
    if n < 0:
        return None
    out = np.empty((n, 3), dtype=np.int32)
    lib().orc_threshold_list(ctypes.byref(g), _p(rho), c, _p(out), n)
    return out


def cluster_crs(crs):
    crs = _i32(crs, (-1, 3))
    label = np.empty(len(crs), dtype=np.int32)
    n = lib().orc_cluster_crs(len(crs), _p(crs), _p(label))
    return label, int(n)


def blob_stats(g, rho, crs, label, nblobs):
    rho = _f32(rho).reshape(-1)
    crs = _i32(crs, (-1, 3))
    label = _i32(label)
    stats = np.empty((nblobs, 8), dtype=np.float64)
    lib().orc_blob_stats(ctypes.byref(g), _p(rho), len(crs), _p(crs), _p(label), nblobs, _p(stats))
    return stats


def full_blobs(g, rho, cutoff):
    """(crs in createFullCrsList order, blob label per voxel, n blobs) or None for cutoff 0."""
    rho = _f32(rho).reshape(-1)
    c = float(np.float32(cutoff))
    if c == 0.0:
        return None
    counts = np.zeros(2, dtype=np.int64)
    cap = 1 << 16
    while True:
        crs = np.empty((cap, 3), dtype=np.int32)
        label = np.empty(cap, dtype=np.int32)
        rc = lib().orc_full_blobs(ctypes.byref(g), _p(rho), c, _p(crs), _p(label), cap, _p(counts))
        if rc == 0:
            return crs[:counts[0]].copy(), label[:counts[0]].copy(), int(counts[1])
        cap = int(counts[0])


def symmetry(g, xyz, rot, shift27, lo, hi):
    xyz = _f64(xyz, (-1, 3))
    rot = _f64(rot, (-1, 12))
    shift27 = _f64(shift27, (27, 3))
    lo, hi = _f64(lo, (3,)), _f64(hi, (3,))
    cap = 27 * len(rot) * len(xyz)
    atom = np.empty(cap, dtype=np.int32)
    image = np.empty(cap, dtype=np.int32)
    out = np.empty((cap, 3), dtype=np.float64)
    n = lib().orc_symmetry(ctypes.byref(g), len(xyz), _p(xyz), len(rot), _p(rot), _p(shift27), _p(lo), _p(hi), _p(atom),
                           _p(image), _p(out), cap)
    return atom[:n].copy(), image[:n].copy(), out[:n].copy()


def nearest(centroid, coords):
    centroid = _f64(centroid, (-1, 3))
    coords = _f64(coords, (-1, 3))
    idx = np.empty(len(centroid), dtype=np.int32)
    dist = np.empty(len(centroid), dtype=np.float64)
    lib().orc_nearest(len(centroid), _p(centroid), len(coords), _p(coords), _p(idx), _p(dist))
    return idx, dist


def sum_abs(rho, cutoff):
    rho = _f32(rho).reshape(-1)
    return lib().orc_sum_abs(_p(rho), len(rho), float(np.float32(cutoff)))


def test_overlap(a, b):
    a, b = _i32(a, (-1, 3)), _i32(b, (-1, 3))
    return bool(lib().orc_test_overlap(len(a), _p(a), len(b), _p(b)))
