"""Drop-in for the reference's operator seam ``pdb_eda.cutils`` / ``pdb_eda.utils`` -- on a B200.

``pdb_eda/ccp4.py:16-19`` and ``pdb_eda/densityAnalysis.py:26-29`` bind ``utils`` to a module exporting thirteen
names; this module exports the same thirteen with the same arguments, return types and float32-narrowing quirks
(the Cython ``float`` parameters), and runs every voxel loop as a CUDA kernel of
``libpdbeda_b200.so``.  ``densityMatrix`` stays duck-typed: anything with ``.header``, ``.density`` and ``.origin``
(this package's :class:`~pdb_eda_b200.ccp4.DensityMatrix` or the reference's own).

Per-call use pays a kernel launch and a device round trip per call; the batched entry points further down
(``sphereSums``, ``sphereLists``, ``fullBlobs`` ...) are what :mod:`pdb_eda_b200.densityAnalysis` uses.
There is no CPU fallback: without the CUDA library these functions raise.
"""
import ctypes
import weakref

import numpy as np
import torch

from . import _device
from . import _lib
from ._device import _ptr, _stream, _as_dev
from ._gc import pausedGC
from ._lib import check

dcutoff = np.sqrt(3)  # adjacency radius of createCrsLists (pdb_eda/cutils.pyx:43)

_foreign = weakref.WeakKeyDictionary()


def _f32(x):
    """The value a Cython ``float`` parameter holds (pdb_eda/cutils.pyx:28,185,205,220,250,273)."""
    return float(np.float32(x))


def deviceMap(densityMatrix):
    """The HBM-resident copy of a duck-typed density matrix."""
    dm = getattr(densityMatrix, "deviceMap", None)
    if dm is not None:
        return dm
    # a foreign object (e.g. the reference's DensityMatrix): cache the upload, revalidate by a cheap fingerprint
    dens = np.asarray(densityMatrix.density)
    finger = (dens.shape, float(dens.sum()), float(np.abs(dens).sum()))
    entry = _foreign.get(densityMatrix)
    if entry is None or entry[0] != finger:
        narrowed = dens.astype(np.float32)
        if not np.array_equal(narrowed.astype(np.float64), dens.astype(np.float64), equal_nan=True):
            raise ValueError("density values are not float32-representable; the device map is float32 like the CCP4 file")
        entry = (finger, _device.DeviceMap.from_host(densityMatrix.header, narrowed, densityMatrix.origin))
        _foreign[densityMatrix] = entry
    return entry[1]


def _xyz64(xyzCoord):
    return np.asarray(xyzCoord, dtype=np.float64).reshape(-1, 3)  # float32 widens exactly


# ================================================================================================ the 13 names
def testOverlap(selfBlob, otherBlob):
    """True iff some voxel of one blob is identical or 26-adjacent to a voxel of the other (pdb_eda/cutils.pyx:8-25)."""
    a = _blob_array(selfBlob)
    b = _blob_array(otherBlob)
    if len(a) == 0 or len(b) == 0:
        return False
    crs = np.concatenate((a, b))
    owner = np.concatenate((np.zeros(len(a), np.int32), np.ones(len(b), np.int32)))
    return len(overlapPairs(crs, owner)) > 0


def sumOfAbs(array, cutoff):
    """Sum of |v| over values with |v| > float32(cutoff) (pdb_eda/cutils.pyx:28-39)."""
    if hasattr(array, "header") and hasattr(array, "density"):
        return deviceMap(array).sum_abs(cutoff)
    _device.require_cuda()
    lib = _lib.load()
    vals = array if isinstance(array, np.ndarray) else np.asarray(array, dtype=np.float64)
    if vals.size == 0:
        return 0
    out = torch.empty(1, dtype=torch.float64, device="cuda")
    ws = torch.empty(int(lib.pe_stats_workspace_bytes()), dtype=torch.uint8, device="cuda")
    if vals.dtype == np.float32:
        t = _as_dev(vals.reshape(-1), torch.float32, "cuda")
        check(lib.pe_map_sum_abs(_ptr(t), t.numel(), ctypes.c_float(_f32(cutoff)), _ptr(out), _ptr(ws), _stream()), "pe_map_sum_abs")
    else:
        t = _as_dev(vals.reshape(-1), torch.float64, "cuda")
        check(lib.pe_sum_abs_f64(_ptr(t), t.numel(), ctypes.c_float(_f32(cutoff)), _ptr(out), _ptr(ws), _stream()), "pe_sum_abs_f64")
    return out.item()


@pausedGC
def createCrsLists(crsList):
    """Disjoint 26-connected voxel lists, in the reference's creation order (pdb_eda/cutils.pyx:44-70)."""
    crsList = list(crsList)
    if not crsList:
        return []
    label, n = _device.cluster_crs(np.asarray(crsList, dtype=np.int32).reshape(-1, 3))
    label = label.cpu().numpy()
    out = [[] for _ in range(n)]
    for crs, lab in zip(crsList, label.tolist()):
        out[lab].append(crs)
    return out


class SymAtom:
    """An atom image: delegates to the wrapped atom except for ``coord`` and ``symmetry`` (pdb_eda/cutils.pyx:105-123)."""

    def __init__(self, atom, coord, symmetry):
        self.atom = atom
        self.coord = coord
        self.symmetry = symmetry

    def __getattr__(self, attr):
        return getattr(self.atom, attr)


@pausedGC
def createSymmetryAtoms(atomList, rotationMats, orthoMat, xs, ys, zs):
    """All symmetry images inside the map's circumscribed box +- 5 A (pdb_eda/cutils.pyx:73-103)."""
    atomIndex, image, coords = symmetryImages([atom.coord for atom in atomList], rotationMats, orthoMat, xs, ys, zs)
    nops = len(rotationMats)
    symmetryOf = {}                                   # image code -> (i, j, k, operator); 27 x nops distinct codes at most
    out = []
    for a, code, xyz in zip(atomIndex.tolist(), image.tolist(), coords):
        symmetry = symmetryOf.get(code)
        if symmetry is None:
            img, op = divmod(code, nops)
            symmetry = symmetryOf[code] = (img // 9 - 1, (img // 3) % 3 - 1, img % 3 - 1, op)
        atom = atomList[a]
        out.append(SymAtom(atom, atom.coord if symmetry == (0, 0, 0, 0) else xyz, symmetry))
    return out


def getPointDensityFromCrs(densityMatrix, crsCoord):
    """Periodic-wrapped density lookup; 0 where the cell is not covered (pdb_eda/cutils.pyx:125-145)."""
    crs = np.asarray(list(crsCoord), dtype=np.int64).reshape(1, 3)
    val, ok = deviceMap(densityMatrix).point_density(crs.astype(np.int32))
    if not bool(ok.item()):
        return 0
    return np.float64(val.item())


def testValidCrs(densityMatrix, crsCoord):
    """Is the (wrapped) index covered by the stored map (pdb_eda/cutils.pyx:147-167)."""
    crs = np.asarray(list(crsCoord), dtype=np.int32).reshape(1, 3)
    return bool(deviceMap(densityMatrix).point_density(crs)[1].item())


def testValidCrsList(densityMatrix, crsList):
    """All indices valid (pdb_eda/cutils.pyx:169-183)."""
    crs = np.asarray(list(crsList), dtype=np.int32).reshape(-1, 3)
    if len(crs) == 0:
        return True
    return bool(deviceMap(densityMatrix).point_density(crs)[1].all().item())


def createFullCrsList(densityMatrix, cutoff):
    """Voxels of the unique sub-volume with rho >= cutoff (> 0) or rho <= cutoff (< 0), column slowest
    (pdb_eda/cutils.pyx:185-203); None for cutoff 0."""
    c32 = _f32(cutoff)
    if c32 == 0.0:
        return None
    res = deviceMap(densityMatrix).blob_label(c32 if c32 > 0 else 0.0, c32 if c32 < 0 else 0.0)
    part = res[0] if c32 > 0 else res[1]
    return list(map(tuple, part["crs"].cpu().numpy().tolist()))


def getSphereCrsFromXyz(densityMatrix, xyzCoord, radius, densityCutoff=0):
    """Voxels within ``radius`` of a point that pass the density predicate, in box order (pdb_eda/cutils.pyx:220-248)."""
    res = deviceMap(densityMatrix).sphere_lists(_xyz64(xyzCoord), [_f32(radius)], _f32(densityCutoff))
    return list(map(tuple, res["crs"].cpu().numpy().tolist()))


def getSphereCrsFromXyzList(densityMatrix, xyzCoordList, radius, densityCutoff=0):
    """Set-union of the spheres of several points; ``radius`` may be a list (pdb_eda/cutils.pyx:250-271)."""
    xyz = _xyz64(xyzCoordList)
    if isinstance(radius, list):
        n = min(len(xyz), len(radius))
        xyz, radii = xyz[:n], [_f32(r) for r in radius[:n]]
    else:
        radii = [_f32(radius)] * len(xyz)
    if len(xyz) == 0:
        return set()
    res = deviceMap(densityMatrix).sphere_lists(xyz, radii, _f32(densityCutoff))
    return set(map(tuple, res["crs"].cpu().numpy().tolist()))


def testValidXyz(densityMatrix, xyzCoord, radius):
    """Every in-sphere voxel lies inside the stored map (pdb_eda/cutils.pyx:273-294)."""
    return testValidXyzList(densityMatrix, [xyzCoord], radius)


def testValidXyzList(densityMatrix, xyzCoordList, radius):
    """testValidXyz for every point (pdb_eda/cutils.pyx:296-313)."""
    xyz = _xyz64(xyzCoordList)
    if len(xyz) == 0:
        return True
    out = deviceMap(densityMatrix).sphere_sums(xyz, [_f32(radius)] * len(xyz))
    return bool((out[:, 6] != 0).all().item())


# ================================================================================================ batched forms
def getTotalDensityFromXyz(densityMatrix, xyzCoord, radius, densityCutoff=0):
    """DensityMatrix.getTotalDensityFromXyz (pdb_eda/ccp4.py:418-435) without materialising the voxel list."""
    c32 = _f32(densityCutoff)
    out = deviceMap(densityMatrix).sphere_sums(_xyz64(xyzCoord), [_f32(radius)], None, max(c32, 0.0), min(c32, 0.0))
    row = out[0].tolist()
    return row[1] if c32 == 0.0 else (row[3] if c32 > 0 else row[5])


def sphereSums(densityMatrix, xyzCoordList, radius, groupStart=None, positiveCutoff=0.0, negativeCutoff=0.0):
    """Per atom (or per group of atoms, set-union) sphere sums as an (n, 8) float64 array; columns as documented
    for ``pe_sphere_sums`` in include/pdbeda_b200.h."""
    xyz = _xyz64(xyzCoordList)
    radii = np.broadcast_to(np.asarray(radius, dtype=np.float32), (len(xyz),)) if np.ndim(radius) == 0 else np.asarray(radius, dtype=np.float32)
    return deviceMap(densityMatrix).sphere_sums(xyz, np.ascontiguousarray(radii), groupStart, _f32(positiveCutoff),
                                               _f32(negativeCutoff)).cpu().numpy()


def sphereLists(densityMatrix, xyzCoordList, radius, densityCutoff=0.0, values=False, labels=False):
    """getSphereCrsFromXyz for many atoms at once; see :meth:`DeviceMap.sphere_lists`."""
    xyz = _xyz64(xyzCoordList)
    radii = np.broadcast_to(np.asarray(radius, dtype=np.float32), (len(xyz),)) if np.ndim(radius) == 0 else np.asarray(radius, dtype=np.float32)
    return deviceMap(densityMatrix).sphere_lists(xyz, np.ascontiguousarray(radii), _f32(densityCutoff), values, labels)


def _blob_array(blob):
    arr = getattr(blob, "crsArray", None)
    if arr is None:
        arr = np.asarray(list(blob.crsList), dtype=np.int32).reshape(-1, 3)
    return arr


def _make_blobs(densityMatrix, crs, label, stats):
    """DensityBlob objects from per-voxel labels and the per-blob sums (pdb_eda/ccp4.py:534-545)."""
    from .ccp4 import DensityBlob
    n = stats[:, 0]
    total = stats[:, 1]
    with np.errstate(divide="ignore", invalid="ignore"):
        centroid = stats[:, 2:5] / total[:, None]
        center = stats[:, 5:8] / n[:, None]
    unit = densityMatrix.header.unitVolume
    order = np.argsort(label, kind="stable")
    bounds = np.searchsorted(label[order], np.arange(len(stats) + 1)).tolist()
    grouped = crs[order]                              # one gather; every blob's members are a slice (a view) of it
    centroids, centers, totals = centroid.tolist(), center.tolist(), total.tolist()
    return [DensityBlob(centroids[b], centers[b], totals[b], unit * (bounds[b + 1] - bounds[b]), grouped[bounds[b]:bounds[b + 1]],
                        densityMatrix) for b in range(len(stats))]


@pausedGC
def fullBlobs(densityMatrix, positiveCutoff, negativeCutoff):
    """[green, red] = createFullBlobList(+c), createFullBlobList(-c) from one pass over the map
    (pdb_eda/ccp4.py:463-485, pdb_eda/densityAnalysis.py:392-412); None for a zero cutoff."""
    res = deviceMap(densityMatrix).blob_label(max(_f32(positiveCutoff), 0.0), min(_f32(negativeCutoff), 0.0))
    out = []
    for part in res:
        if part is None:
            out.append(None)
            continue
        out.append(_make_blobs(densityMatrix, part["crs"].cpu().numpy(), part["label"].cpu().numpy(),
                               part["stats"].cpu().numpy()))
    return out


def crsStats(densityMatrix, crs, label=None, take=None, nClusters=1):
    """Per-cluster sums n, sum rho, sum rho*xyz, sum xyz of a labelled voxel list -> (nClusters, 8) device tensor."""
    dmap = deviceMap(densityMatrix)
    lib = dmap.lib
    crs = _as_dev(crs, torch.int32, dmap.device, (-1, 3))
    label_t = _as_dev(label, torch.int32, dmap.device, (-1,)) if label is not None else None
    take_t = None
    if take is not None:
        take_t = take.to(device=dmap.device, dtype=torch.uint8).contiguous() if isinstance(take, torch.Tensor) else \
            torch.from_numpy(np.ascontiguousarray(np.asarray(take, dtype=np.uint8))).to(dmap.device)
    stats = torch.empty((int(nClusters), 8), dtype=torch.float64, device=dmap.device)
    check(lib.pe_crs_stats(ctypes.byref(dmap.geom), _ptr(dmap.rho), crs.shape[0], _ptr(crs), _ptr(label_t), _ptr(take_t),
                           int(nClusters), _ptr(stats), _stream()), "pe_crs_stats")
    return stats


def pairMetrics(foMatrix, diffMatrix, crs, label=None, take=None, nGroups=1):
    """Per-group RSCC / RSR sums on the Fo map and the on-the-fly Fc map (``pe_pair_metrics``) -> (nGroups, 8) tensor."""
    fo, diff = deviceMap(foMatrix), deviceMap(diffMatrix)
    if fo.rho.numel() != diff.rho.numel():
        raise ValueError("the 2Fo-Fc and Fo-Fc maps must share one grid")
    crs = _as_dev(crs, torch.int32, fo.device, (-1, 3))
    label_t = _as_dev(label, torch.int32, fo.device, (-1,)) if label is not None else None
    take_t = None
    if take is not None:
        take_t = take.to(device=fo.device, dtype=torch.uint8).contiguous() if isinstance(take, torch.Tensor) else \
            torch.from_numpy(np.ascontiguousarray(np.asarray(take, dtype=np.uint8))).to(fo.device)
    out = torch.empty((int(nGroups), 8), dtype=torch.float64, device=fo.device)
    check(fo.lib.pe_pair_metrics(ctypes.byref(fo.geom), _ptr(fo.rho), _ptr(diff.rho), crs.shape[0], _ptr(crs), _ptr(label_t),
                                 _ptr(take_t), int(nGroups), _ptr(out), _stream()), "pe_pair_metrics")
    return out


@pausedGC
def blobsFromCrsList(densityMatrix, crsList):
    """createBlobList (pdb_eda/ccp4.py:475-485): cluster + per-blob sums on the device."""
    crs = np.asarray(list(crsList), dtype=np.int32).reshape(-1, 3)
    if len(crs) == 0:
        return []
    label, n = _device.cluster_crs(crs)
    stats = crsStats(densityMatrix, crs, label, None, n).cpu().numpy()
    return _make_blobs(densityMatrix, crs, label.cpu().numpy(), stats)


def blobFromCrsList(densityMatrix, crsList):
    """DensityBlob.fromCrsList (pdb_eda/ccp4.py:522-545)."""
    from .ccp4 import DensityBlob
    crs = np.asarray(list(crsList), dtype=np.int32).reshape(-1, 3)
    st = crsStats(densityMatrix, crs, None, None, 1).cpu().numpy()[0]
    n = len(crs)
    with np.errstate(divide="ignore", invalid="ignore"):
        centroid = (st[2:5] / st[1]).tolist()
        center = (st[5:8] / n).tolist() if n else [float("nan")] * 3
    return DensityBlob(centroid, center, float(st[1]), densityMatrix.header.unitVolume * n, crs, densityMatrix)


def clusterCrs(crs, group=None, wantFirst=False, device="cuda"):
    """Grouped 26-connected clustering of a voxel list with repeats -> (label, first or None, n clusters)."""
    _device.require_cuda()
    lib = _lib.load()
    crs = _as_dev(crs, torch.int32, device, (-1, 3))
    n = crs.shape[0]
    group_t = _as_dev(group, torch.int32, crs.device, (-1,)) if group is not None else None
    label = torch.empty(n, dtype=torch.int32, device=crs.device)
    first = torch.empty(n, dtype=torch.uint8, device=crs.device) if wantFirst else None
    counts = torch.zeros(2, dtype=torch.int64, device=crs.device)
    ws = torch.empty(max(int(lib.pe_cluster_workspace_bytes(n)), 256), dtype=torch.uint8, device=crs.device)
    check(lib.pe_cluster_crs_grouped(n, _ptr(crs), _ptr(group_t), _ptr(label), _ptr(first), _ptr(counts), _ptr(ws), _stream()),
          "pe_cluster_crs_grouped")
    ncl, bad = counts.tolist()
    if bad:
        raise _lib.PdbEdaLibError("clusterCrs: voxel index or group id outside the supported key range")
    return label, first, int(ncl)


def overlapPairs(crs, owner, group=None, device="cuda"):
    """All overlapping (owner_a < owner_b) pairs of a voxel list, sorted -> (m, 2) int32 numpy array."""
    _device.require_cuda()
    lib = _lib.load()
    crs = _as_dev(crs, torch.int32, device, (-1, 3))
    n = crs.shape[0]
    owner_t = _as_dev(owner, torch.int32, crs.device, (-1,))
    group_t = _as_dev(group, torch.int32, crs.device, (-1,)) if group is not None else None
    cap = max(1024, n // 4)
    while True:
        counts = torch.zeros(2, dtype=torch.int64, device=crs.device)
        pairs = torch.empty((cap, 2), dtype=torch.int32, device=crs.device)
        ws = torch.empty(max(int(lib.pe_overlap_workspace_bytes(n, cap)), 256), dtype=torch.uint8, device=crs.device)
        check(lib.pe_overlap_pairs(n, _ptr(crs), _ptr(owner_t), _ptr(group_t), cap, _ptr(counts), _ptr(pairs), _ptr(ws),
                                   _stream()), "pe_overlap_pairs")
        found, bad = counts.tolist()
        if bad:
            raise _lib.PdbEdaLibError("overlapPairs: voxel index or group id outside the supported key range")
        if found <= cap:
            break
        cap = int(found) * 2
    out = pairs[:found].cpu().numpy()
    if len(out):
        out = out[np.lexsort((out[:, 1], out[:, 0]))]
    return out


def symmetryImages(coords, rotationMats, orthoMat, xs, ys, zs, densityMatrix=None):
    """Device part of createSymmetryAtoms: (atom index, image code, float64 xyz) numpy arrays in the reference's
    order.  Image code = ((i+1)*9 + (j+1)*3 + (k+1)) * nops + op."""
    _device.require_cuda()
    xyz = _xyz64(coords)
    rot = np.asarray(rotationMats, dtype=np.float64).reshape(-1, 12)
    # lattice shifts exactly as the reference forms them: np.dot(orthoMat, (i, j, k)) (pdb_eda/cutils.pyx:98)
    shift = np.array([np.dot(orthoMat, (i, j, k)) for i in (-1, 0, 1) for j in (-1, 0, 1) for k in (-1, 0, 1)], dtype=np.float64)
    lo = [xs[0] - 5, ys[0] - 5, zs[0] - 5]
    hi = [xs[-1] + 5, ys[-1] + 5, zs[-1] + 5]
    helper = _SymmetryHelper.get()
    atom, image, out = helper.symmetry_expand(xyz, rot, shift, lo, hi)
    return atom.cpu().numpy(), image.cpu().numpy(), out.cpu().numpy()


class _SymmetryHelper:
    """pe_symmetry_expand only needs the BLAS accumulation order of pe_geom; a 1-voxel dummy map carries it."""
    _inst = None

    @classmethod
    def get(cls):
        if cls._inst is None:
            g = _lib.PeGeom()
            perm, fma = _device._blas.probe()
            for a in range(3):
                g.ncrs[a] = g.crs_interval[a] = g.xyz_interval[a] = g.unique_ncrs[a] = 1
                g.map2xyz[a] = g.map2crs[a] = a
                g.mv_perm[a] = perm[a]
                g.grid_length[a] = 1.0
            g.mv_fma = fma
            g.orthogonal = 1
            for k in (0, 4, 8):
                g.ortho[k] = g.deortho[k] = 1.0
            cls._inst = _device.DeviceMap(g, torch.zeros(1, dtype=torch.float32, device="cuda"))
        return cls._inst
