"""Device side of the voxel path: torch owns the buffers, ``libpdbeda_b200.so`` (C ABI, ctypes) does the work.

PyTorch is plumbing here (device memory, streams); every computation is a hand-written sm_100a kernel behind the
entry points declared in ``include/pdbeda_b200.h``.  There is no CPU fallback: without the library or without a
CUDA device every call raises.
"""
import ctypes

import numpy as np
import torch

from . import _blas
from . import _lib
from ._lib import PE_SPHERE_NOUT, PdbEdaLibError, PeGeom, check


def geom_from_header(header, origin=None):
    """Packs a (duck-typed) ``DensityHeader`` into ``pe_geom``.

    Reads exactly the attributes the reference's ``cutils`` reads from ``densityMatrix.header``
    (pdb_eda/ccp4.py:158-286), so it accepts this package's header and the reference's alike.
    """
    g = PeGeom()
    perm, fma = _blas.probe()
    origin = header.origin if origin is None else origin
    for a in range(3):
        g.ncrs[a] = int(header.ncrs[a])
        g.crs_start[a] = int(header.crsStart[a])
        g.xyz_interval[a] = int(header.xyzInterval[a])
        g.crs_interval[a] = int(header.crsInterval[a])
        g.unique_ncrs[a] = int(header.uniqueNcrs[a])
        g.map2xyz[a] = int(header.map2xyz[a])
        g.map2crs[a] = int(header.map2crs[a])
        g.mv_perm[a] = perm[a]
        g.grid_length[a] = float(header.gridLength[a])
        g.origin[a] = float(origin[a])
    g.mv_fma = fma
    g.orthogonal = 1 if (header.alpha == header.beta == header.gamma == 90) else 0
    ortho = np.asarray(header.orthoMat, dtype=np.float64).reshape(9)
    deortho = np.asarray(header.deOrthoMat, dtype=np.float64).reshape(9)
    for k in range(9):
        g.ortho[k] = float(ortho[k])
        g.deortho[k] = float(deortho[k])
    return g


def _stream():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def _ptr(t):
    return ctypes.c_void_p(t.data_ptr()) if t is not None else ctypes.c_void_p(0)


def require_cuda():
    if not torch.cuda.is_available():
        raise PdbEdaLibError("pdb_eda_b200 needs a CUDA device (B200, sm_100a); there is no CPU fallback")
    return _lib.require_device()


class Workspace:
    """A grow-only scratch buffer so that no allocation happens inside timed regions once it is warm."""

    def __init__(self, device):
        self.device = device
        self.buf = None

    def get(self, nbytes):
        nbytes = max(int(nbytes), 256)
        if self.buf is None or self.buf.numel() < nbytes:
            self.buf = torch.empty(nbytes, dtype=torch.uint8, device=self.device)
        return self.buf


def _as_dev(x, dtype, device, shape=None):
    """Host array-like or torch tensor -> contiguous device tensor of ``dtype`` (exact widening only)."""
    if isinstance(x, torch.Tensor):
        t = x.to(device=device, dtype=dtype).contiguous()
    else:
        npdt = {torch.float64: np.float64, torch.float32: np.float32, torch.int32: np.int32, torch.int64: np.int64}[dtype]
        t = torch.from_numpy(np.ascontiguousarray(np.asarray(x, dtype=npdt))).to(device)
    if shape is not None:
        t = t.reshape(shape)
    return t


class DeviceMap:
    """One CCP4 map resident in HBM: float32 ``rho[section][row][column]`` plus its ``pe_geom``.

    Mirrors what the reference keeps per ``DensityMatrix`` (pdb_eda/ccp4.py:319-341): header geometry and the
    voxel array.  All methods enqueue on torch's current stream and return torch tensors unless noted.
    """

    def __init__(self, geom, rho):
        require_cuda()
        if not (isinstance(rho, torch.Tensor) and rho.is_cuda and rho.dtype == torch.float32 and rho.is_contiguous()):
            raise PdbEdaLibError("DeviceMap needs a contiguous float32 CUDA tensor")
        n = geom.ncrs[0] * geom.ncrs[1] * geom.ncrs[2]
        if rho.numel() != n:
            raise PdbEdaLibError("voxel tensor has %d elements, header says %d" % (rho.numel(), n))
        self.geom = geom
        self.rho = rho
        self.device = rho.device
        self.lib = _lib.load()
        self.ws = Workspace(self.device)
        self._mean_std = None
        self._sum_abs = {}

    @classmethod
    def from_host(cls, header, values, origin=None, device="cuda", pinned=None):
        """``values``: anything numpy can view as the flat float32 payload (column fastest).  ``pinned``: the page-locked
        uint8 tensor that backs ``values`` (set by ``ccp4.parse``): the upload is then a single asynchronous DMA."""
        require_cuda()
        if pinned is not None:
            rho = pinned.view(torch.float32).to(device, non_blocking=True)
        else:
            arr = np.ascontiguousarray(np.asarray(values, dtype=np.float32)).reshape(-1)
            if not arr.flags.writeable:
                arr = arr.copy()
            rho = torch.from_numpy(arr).to(device)
        return cls(geom_from_header(header, origin), rho)

    # ------------------------------------------------------------------------------------------ whole-map sums
    @property
    def n_voxels(self):
        return self.rho.numel()

    def mean_std(self):
        """(meanDensity, stdDensity) of pdb_eda/ccp4.py:343-363 as Python floats (synchronises)."""
        if self._mean_std is None:
            out = torch.empty(2, dtype=torch.float64, device=self.device)
            ws = self.ws.get(self.lib.pe_stats_workspace_bytes())
            check(self.lib.pe_map_mean_std(_ptr(self.rho), self.n_voxels, _ptr(out), _ptr(ws), _stream()), "pe_map_mean_std")
            m, s = out.tolist()
            self._mean_std = (m, s)
        return self._mean_std

    def sum_abs(self, cutoff):
        """sumOfAbs(densityArray, cutoff) (pdb_eda/cutils.pyx:28-39); cutoff is narrowed to float32 like Cython does."""
        c32 = float(np.float32(cutoff))
        if c32 not in self._sum_abs:
            out = torch.empty(1, dtype=torch.float64, device=self.device)
            ws = self.ws.get(self.lib.pe_stats_workspace_bytes())
            check(self.lib.pe_map_sum_abs(_ptr(self.rho), self.n_voxels, ctypes.c_float(c32), _ptr(out), _ptr(ws), _stream()),
                  "pe_map_sum_abs")
            self._sum_abs[c32] = out.item()
        return self._sum_abs[c32]

    def invalidate(self):
        self._mean_std = None
        self._sum_abs = {}

    # ------------------------------------------------------------------------------------------ point conversions
    def point_density(self, crs):
        """(density float32, valid uint8) for n x 3 un-wrapped indices (pdb_eda/cutils.pyx:125-167)."""
        crs = _as_dev(crs, torch.int32, self.device, (-1, 3))
        n = crs.shape[0]
        out = torch.empty(n, dtype=torch.float32, device=self.device)
        valid = torch.empty(n, dtype=torch.uint8, device=self.device)
        check(self.lib.pe_point_density(ctypes.byref(self.geom), _ptr(self.rho), n, _ptr(crs), _ptr(out), _ptr(valid), _stream()),
              "pe_point_density")
        return out, valid

    def xyz2crs(self, xyz):
        xyz = _as_dev(xyz, torch.float64, self.device, (-1, 3))
        out = torch.empty(xyz.shape, dtype=torch.int32, device=self.device)
        check(self.lib.pe_xyz2crs(ctypes.byref(self.geom), xyz.shape[0], _ptr(xyz), _ptr(out), _stream()), "pe_xyz2crs")
        return out

    def crs2xyz(self, crs):
        crs = _as_dev(crs, torch.int32, self.device, (-1, 3))
        out = torch.empty(crs.shape, dtype=torch.float64, device=self.device)
        check(self.lib.pe_crs2xyz(ctypes.byref(self.geom), crs.shape[0], _ptr(crs), _ptr(out), _stream()), "pe_crs2xyz")
        return out

    # ------------------------------------------------------------------------------------------ atom spheres
    def sphere_sums(self, xyz, radius, group_start=None, cut_pos=0.0, cut_neg=0.0, out=None):
        """Per atom (or per group of atoms, set-union) the PE_SPHERE_NOUT sums documented in the header."""
        xyz = _as_dev(xyz, torch.float64, self.device, (-1, 3))
        n = xyz.shape[0]
        radius = _as_dev(radius, torch.float32, self.device, (-1,))
        if radius.numel() != n:
            raise PdbEdaLibError("sphere_sums: %d radii for %d atoms" % (radius.numel(), n))
        if group_start is not None:
            group_start = _as_dev(group_start, torch.int32, self.device, (-1,))
            n_groups = group_start.numel() - 1
        else:
            n_groups = n
        if out is None:
            out = torch.empty((n_groups, PE_SPHERE_NOUT), dtype=torch.float64, device=self.device)
        ws = self.ws.get(self.lib.pe_sphere_workspace_bytes(n))
        check(self.lib.pe_sphere_sums(ctypes.byref(self.geom), _ptr(self.rho), n, _ptr(xyz), _ptr(radius), n_groups,
                                      _ptr(group_start), ctypes.c_float(float(np.float32(cut_pos))),
                                      ctypes.c_float(float(np.float32(cut_neg))), _ptr(out), _ptr(ws), _stream()),
              "pe_sphere_sums")
        return out

    def sphere_lists(self, xyz, radius, cutoff=0.0, want_values=False, want_labels=False):
        """getSphereCrsFromXyz for a batch of atoms (pdb_eda/cutils.pyx:220-248).

        Returns a dict of device tensors: ``count`` (n), ``offset`` (n+1, int64), ``box`` (n x 6: low corner and
        extents), ``crs`` (total x 3 un-wrapped indices in the reference's order), ``atom`` (total: owning atom),
        and optionally ``value`` (wrapped density) and ``label`` (cluster number inside the atom's list, in
        createCrsLists order).
        """
        xyz = _as_dev(xyz, torch.float64, self.device, (-1, 3))
        n = xyz.shape[0]
        radius = _as_dev(radius, torch.float32, self.device, (-1,))
        c32 = ctypes.c_float(float(np.float32(cutoff)))
        count = torch.empty(n, dtype=torch.int32, device=self.device)
        box = torch.empty((n, 6), dtype=torch.int32, device=self.device)
        g = ctypes.byref(self.geom)
        check(self.lib.pe_sphere_count(g, _ptr(self.rho), n, _ptr(xyz), _ptr(radius), c32, _ptr(count), _ptr(box), _stream()),
              "pe_sphere_count")
        offset = torch.zeros(n + 1, dtype=torch.int64, device=self.device)
        torch.cumsum(count, 0, out=offset[1:])
        if n:
            total = int(offset[-1].item())
            max_box = int((box[:, 3].long() * box[:, 4].long() * box[:, 5].long()).max().item())
        else:
            total, max_box = 0, 0
        index = torch.empty(total, dtype=torch.int32, device=self.device)
        value = torch.empty(total, dtype=torch.float32, device=self.device) if want_values else None
        label = torch.empty(total, dtype=torch.int32, device=self.device) if want_labels else None
        if n and max_box > 0:
            check(self.lib.pe_sphere_fill(g, _ptr(self.rho), n, _ptr(xyz), _ptr(radius), c32, _ptr(offset), max_box,
                                          _ptr(index), _ptr(value), _ptr(label), _stream()), "pe_sphere_fill")
        atom = torch.repeat_interleave(torch.arange(n, device=self.device), count.long())
        b = box[atom]
        p = index.long()
        nrs = b[:, 4].long() * b[:, 5].long()
        ic = p // nrs.clamp(min=1)
        ir = (p // b[:, 5].long().clamp(min=1)) % b[:, 4].long().clamp(min=1)
        is_ = p % b[:, 5].long().clamp(min=1)
        crs = torch.stack((b[:, 0].long() + ic, b[:, 1].long() + ir, b[:, 2].long() + is_), dim=1).to(torch.int32)
        res = {"count": count, "offset": offset, "box": box, "crs": crs, "atom": atom.to(torch.int32)}
        if want_values:
            res["value"] = value
        if want_labels:
            res["label"] = label
        return res

    # ------------------------------------------------------------------------------------------ difference-map blobs
    def blob_label(self, cut_pos, cut_neg, cap_voxels=None, cap_blobs=None):
        """createFullBlobList(+cut_pos) and createFullBlobList(cut_neg) in one pass (pdb_eda/ccp4.py:463-485).

        Returns a list of two dicts (green, red) or None entries for skipped classes: ``crs`` (n x 3, the
        createFullCrsList order), ``value``, ``label`` (blob number in createCrsLists order), ``stats``
        (n_blobs x 8: n, sum rho, sum rho*xyz, sum xyz).  Synchronises to read the counts; retries with larger
        capacities on overflow.
        """
        g = self.geom
        u = [g.unique_ncrs[0], g.unique_ncrs[1], g.unique_ncrs[2]]
        nvox = u[0] * u[1] * u[2]
        cp = float(np.float32(cut_pos))
        cn = float(np.float32(cut_neg))
        if cap_voxels is None:
            cap_voxels = max(4096, nvox // 64)
        if cap_blobs is None:
            cap_blobs = max(1024, cap_voxels // 4)
        while True:
            cap_voxels = int(min(cap_voxels, max(nvox, 1)))
            cap_blobs = int(min(cap_blobs, cap_voxels))
            counts = torch.empty(5, dtype=torch.int64, device=self.device)
            key = torch.empty(2 * cap_voxels, dtype=torch.int32, device=self.device)
            value = torch.empty(2 * cap_voxels, dtype=torch.float32, device=self.device)
            label = torch.empty(2 * cap_voxels, dtype=torch.int32, device=self.device)
            stats = torch.empty((2 * cap_blobs, 8), dtype=torch.float64, device=self.device)
            ws = self.ws.get(self.lib.pe_blob_workspace_bytes(ctypes.byref(g), cap_voxels))
            check(self.lib.pe_blob_label(ctypes.byref(g), _ptr(self.rho), ctypes.c_float(cp), ctypes.c_float(cn), cap_voxels,
                                         cap_blobs, _ptr(counts), _ptr(key), _ptr(value), _ptr(label), _ptr(stats), _ptr(ws),
                                         _stream()), "pe_blob_label")
            c = counts.tolist()
            if c[4] == 0:
                break
            if c[4] == 2:
                raise PdbEdaLibError("pe_blob_label: the sparse kernel's blocks were not spread evenly over the SMs; no result")
            need_v = max(c[0], c[2])
            need_b = max(c[1], c[3], 1)
            if need_v > cap_voxels:
                cap_voxels = need_v
                cap_blobs = max(cap_blobs, need_v // 4)
            else:
                cap_blobs = max(need_b, cap_blobs * 4)
        out = []
        for k, used in ((0, cp > 0.0), (1, cn < 0.0)):
            if not used:
                out.append(None)
                continue
            nfg, nb = c[2 * k], c[2 * k + 1]
            kk = key[k * cap_voxels:k * cap_voxels + nfg].long() & 0xFFFFFFFF
            crs = torch.stack((kk // (u[1] * u[2]), (kk // u[2]) % u[1], kk % u[2]), dim=1).to(torch.int32)
            out.append({"n_voxels": nfg, "n_blobs": nb, "crs": crs,
                        "value": value[k * cap_voxels:k * cap_voxels + nfg],
                        "label": label[k * cap_voxels:k * cap_voxels + nfg],
                        "stats": stats[k * cap_blobs:k * cap_blobs + nb]})
        return out

    # ------------------------------------------------------------------------------------------ symmetry / distances
    def symmetry_expand(self, xyz, rot, shift27, lo, hi):
        """createSymmetryAtoms (pdb_eda/cutils.pyx:73-103): returns (atom index, image code, xyz) device tensors."""
        xyz = _as_dev(xyz, torch.float64, self.device, (-1, 3))
        rot = _as_dev(rot, torch.float64, self.device, (-1, 12))
        shift = _as_dev(shift27, torch.float64, self.device, (27, 3))
        n, nops = xyz.shape[0], rot.shape[0]
        lo_c = (ctypes.c_double * 3)(*[float(v) for v in lo])
        hi_c = (ctypes.c_double * 3)(*[float(v) for v in hi])
        ws = self.ws.get(self.lib.pe_symmetry_workspace_bytes(n, nops))
        count = torch.zeros(1, dtype=torch.int64, device=self.device)
        cap = max(1024, 2 * n)
        while True:
            atom = torch.empty(cap, dtype=torch.int32, device=self.device)
            image = torch.empty(cap, dtype=torch.int32, device=self.device)
            out = torch.empty((cap, 3), dtype=torch.float64, device=self.device)
            check(self.lib.pe_symmetry_expand(ctypes.byref(self.geom), n, _ptr(xyz), nops, _ptr(rot), _ptr(shift), lo_c, hi_c,
                                              cap, _ptr(count), _ptr(atom), _ptr(image), _ptr(out), _ptr(ws), _stream()),
                  "pe_symmetry_expand")
            kept = int(count.item())
            if kept <= cap:
                return atom[:kept], image[:kept], out[:kept]
            cap = kept


def nearest_atom(centroids, coords, device="cuda"):
    """argmin / min of scipy cdist per centroid (pdb_eda/densityAnalysis.py:932-937) -> (idx int32, dist float64)."""
    require_cuda()
    lib = _lib.load()
    c = _as_dev(centroids, torch.float64, device, (-1, 3))
    a = _as_dev(coords, torch.float64, device, (-1, 3))
    idx = torch.empty(c.shape[0], dtype=torch.int32, device=c.device)
    dist = torch.empty(c.shape[0], dtype=torch.float64, device=c.device)
    check(lib.pe_nearest_atom(c.shape[0], _ptr(c), a.shape[0], _ptr(a), _ptr(idx), _ptr(dist), _stream()), "pe_nearest_atom")
    return idx, dist


def cluster_crs(crs, device="cuda"):
    """createCrsLists (pdb_eda/cutils.pyx:41-70) -> (label per input voxel, number of clusters)."""
    require_cuda()
    lib = _lib.load()
    crs = _as_dev(crs, torch.int32, device, (-1, 3))
    n = crs.shape[0]
    label = torch.empty(n, dtype=torch.int32, device=crs.device)
    ncl = torch.zeros(2, dtype=torch.int64, device=crs.device)
    ws = torch.empty(max(int(lib.pe_cluster_workspace_bytes(n)), 256), dtype=torch.uint8, device=crs.device)
    check(lib.pe_cluster_crs(n, _ptr(crs), _ptr(label), _ptr(ncl), _ptr(ws), _stream()), "pe_cluster_crs")
    count, bad = ncl.tolist()
    if bad:
        raise PdbEdaLibError("cluster_crs: voxel index outside the supported key range (|index| < 2^20)")
    return label, int(count)
