"""CCP4 map object model (host side) -- mirror of ``pdb_eda.ccp4`` for the voxel hot path.

Header and orthogonalisation-matrix parsing stay on the host in Python (BASELINE.json north_star); the voxel
payload lives in HBM as float32 and every voxel loop of the reference (``pdb_eda/cutils.pyx``) runs as a CUDA
kernel through :mod:`pdb_eda_b200.cutils`.  Class, method and attribute names are the reference's
(pdb_eda/ccp4.py:58-594) so that code written against ``pdb_eda.ccp4`` keeps working.
"""
import struct
import threading
import urllib.request

import numpy as np

from . import cutils as utils

urlPrefix = "http://www.ebi.ac.uk/pdbe/coordinates/files/"
urlSuffix = ".ccp4"

# (name, count, struct code) of the first 224 header bytes, in file order (pdb_eda/ccp4.py:149; CCP4 maplib).
_HEADER_WORDS = (
    ("ncrs", 3, "i"), ("mode", 1, "i"), ("crsStart", 3, "i"), ("xyzInterval", 3, "i"), ("cell", 3, "f"),
    ("angles", 3, "f"), ("axisOrder", 3, "i"), ("densityRange", 3, "f"), ("spaceGroup", 1, "i"),
    ("symmetryBytes", 1, "i"), ("skewFlag", 1, "i"), ("skewMat", 9, "f"), ("skewTrans", 3, "f"),
    ("futureUse", 12, "f"), ("originEM", 3, "f"), ("mapChar", 4, "c"), ("machineStamp", 1, "i"), ("rmsd", 1, "f"),
    ("nLabel", 1, "i"),
)
_HEADER_FORMAT = "".join(code * count for _, count, code in _HEADER_WORDS)
HEADER_BYTES = 1024
PINNED_MIN_BYTES = 32 << 20


def readFromPDBID(pdbid, verbose=False):
    """DensityMatrix of a PDB entry's 2Fo-Fc map from PDBe (pdb_eda/ccp4.py:24-35)."""
    return readFromURL(urlPrefix + pdbid.lower() + urlSuffix, pdbid, verbose)


def readFromURL(url, pdbid=None, verbose=False):
    """DensityMatrix from a URL (pdb_eda/ccp4.py:38-55)."""
    with urllib.request.urlopen(url) as handle:
        return parse(handle, pdbid or url, verbose)


def read(ccp4Filename, pdbid=None, verbose=False):
    """DensityMatrix from a .ccp4 file (pdb_eda/ccp4.py:58-74)."""
    with open(ccp4Filename, "rb") as handle:
        return parse(handle, pdbid or ccp4Filename, verbose)


def parse(handle, pdbid, verbose=False):
    """DensityMatrix from a binary handle (pdb_eda/ccp4.py:77-127).

    The payload is viewed with ``np.frombuffer`` instead of being unpacked into one Python float per voxel
    (pdb_eda/ccp4.py:123-124), which is what lets 384^3 and 1024^3 maps load at all.
    """
    header = DensityHeader.fromFileHeader(handle.read(HEADER_BYTES))
    assert header.xlength != 0.0 or header.ylength != 0.0 or header.zlength != 0.0, \
        "Error: Cell dimensions are all 0, Map file will not align with other structures"
    header.symmetry = handle.read(header.symmetryBytes) if header.symmetryBytes > 0 else b""
    if header.endian == "<" and header.mapSize >= PINNED_MIN_BYTES and hasattr(handle, "readinto") and _cudaPresent():
        # the voxels go from the file through a small page-locked ring straight to HBM (reads overlap the DMAs); the host
        # copy that ``densityArray`` / ``density`` expose is fetched back only if somebody asks for it
        rho, got = _streamToDevice(handle, header.mapSize)
        extra = handle.read(1) if got == header.mapSize else b""
        if got != header.mapSize or extra:
            raise AssertionError("Error: file holds %s map bytes, the header promises %d" % ("more than %d" % got if extra else got, header.mapSize))
        return DensityMatrix._fromDevice(header, header.origin, rho.view(_torch().float32), pdbid)
    payload = handle.read()
    if len(payload) != header.mapSize:
        # the reference's size assertions all fail once the lengths disagree (pdb_eda/ccp4.py:95-100)
        raise AssertionError("Error: file holds %d bytes after the header, expected %d symmetry + %d map bytes"
                             % (len(payload) + header.symmetryBytes, header.symmetryBytes, header.mapSize))
    voxels = np.frombuffer(payload, dtype=np.dtype(header.endian + "f4"))
    return DensityMatrix(header, header.origin, voxels, pdbid)


def _torch():
    import torch
    return torch


def _cudaPresent():
    try:
        return _torch().cuda.is_available()
    except Exception:
        return False


RING_SLOT_BYTES = 8 << 20
_ring = None
_ringLock = threading.Lock()


def _streamToDevice(handle, nbytes):
    """Reads ``nbytes`` from ``handle`` into a new uint8 tensor in HBM through two page-locked slots that live as long as the
    process (page-locking a whole 226 MB map costs ~150 ms on first use, the ring ~30 ms once; profiles/r02f_h2d_load.txt): slot k
    is being filled from the file while slot k-1 is in flight.  Returns (tensor, bytes read)."""
    global _ring
    torch = _torch()
    with _ringLock:
        if _ring is None:
            slots = [torch.empty(RING_SLOT_BYTES, dtype=torch.uint8, pin_memory=True) for _ in range(2)]
            _ring = [[slot, memoryview(slot.numpy()).cast("B"), None] for slot in slots]
        dev = torch.empty(nbytes, dtype=torch.uint8, device="cuda")
        off = k = 0
        while off < nbytes:
            entry = _ring[k & 1]
            if entry[2] is not None:
                entry[2].synchronize()               # the DMA that last read this slot
            got = handle.readinto(entry[1][:min(RING_SLOT_BYTES, nbytes - off)])
            if not got:
                break
            dev[off:off + got].copy_(entry[0][:got], non_blocking=True)
            entry[2] = torch.cuda.Event()
            entry[2].record()
            off += got
            k += 1
    return dev, off


class DensityHeader(object):
    """CCP4 header and the cell geometry derived from it (pdb_eda/ccp4.py:130-316)."""

    @classmethod
    def fromFileHeader(cls, fileHeader):
        mode = int.from_bytes(fileHeader[12:16], byteorder="little")
        endian = "<" if 0 <= mode <= 6 else ">"
        words = struct.unpack(endian + _HEADER_FORMAT, fileHeader[:224])
        labels = fileHeader[224:].replace(b" ", b"")
        return cls(words, labels, endian)

    def __init__(self, headerTuple, labels, endian):
        fields = {}
        pos = 0
        for name, count, _ in _HEADER_WORDS:
            fields[name] = headerTuple[pos:pos + count]
            pos += count
        self.ncrs = fields["ncrs"]
        self.mode = fields["mode"][0]
        self.endian = endian
        self.crsStart = fields["crsStart"]
        self.nintervalX, self.nintervalY, self.nintervalZ = fields["xyzInterval"]
        self.xlength, self.ylength, self.zlength = fields["cell"]
        self.alpha, self.beta, self.gamma = fields["angles"]
        self.col2xyz, self.row2xyz, self.sec2xyz = fields["axisOrder"]
        self.densityMin, self.densityMax, self.densityMean = fields["densityRange"]
        self.spaceGroup = fields["spaceGroup"][0]
        self.symmetryBytes = fields["symmetryBytes"][0]
        self.skewFlag = fields["skewFlag"][0]
        self.skewMat = fields["skewMat"]
        self.skewTrans = fields["skewTrans"]
        self.futureUse = fields["futureUse"]
        self.originEM = fields["originEM"]
        self.mapChar = fields["mapChar"]
        self.machineStamp = fields["machineStamp"][0]
        self.rmsd = fields["rmsd"][0]
        self.nLabel = fields["nLabel"][0]
        self.labels = labels
        self.symmetry = b""

        self.mapSize = 4 * self.ncrs[0] * self.ncrs[1] * self.ncrs[2]
        self.xyzLength = [self.xlength, self.ylength, self.zlength]
        self.xyzInterval = [self.nintervalX, self.nintervalY, self.nintervalZ]
        if min(self.xyzInterval) <= 0:
            raise ValueError("CCP4 header: non-positive sampling interval %r" % (self.xyzInterval,))
        self.map2crs = [self.col2xyz - 1, self.row2xyz - 1, self.sec2xyz - 1]
        if sorted(self.map2crs) != [0, 1, 2]:
            raise ValueError("CCP4 header: MAPC/MAPR/MAPS = %r is not a permutation of 1,2,3" % (fields["axisOrder"],))
        self.gridLength = [length / interval for length, interval in zip(self.xyzLength, self.xyzInterval)]
        self.map2xyz = [self.map2crs.index(axis) for axis in range(3)]
        self.crsInterval = [self.xyzInterval[self.map2crs[a]] for a in range(3)]

        ca, cb, cg = (np.cos(np.pi / 180 * angle) for angle in (self.alpha, self.beta, self.gamma))
        sg = np.sin(np.pi / 180 * self.gamma)
        skew = np.sqrt(1 - ca ** 2 - cb ** 2 - cg ** 2 + 2 * ca * cb * cg)  # cell volume / (a b c)
        self.unitVolume = self.xlength * self.ylength * self.zlength / self.nintervalX / self.nintervalY / self.nintervalZ * skew
        # Orthogonalisation matrix, 'Biomolecular Crystallography' (Rupp) p. 233 (pdb_eda/ccp4.py:248-250)
        self.orthoMat = [[self.xlength, self.ylength * cg, self.zlength * cb],
                         [0, self.ylength * sg, self.zlength * (ca - cb * cg) / sg],
                         [0, 0, self.zlength * skew / sg]]
        self.deOrthoMat = np.linalg.inv(self.orthoMat)
        self.deOrthoMat[abs(self.deOrthoMat) < 1e-10] = 0.0
        self.origin = self._calculateOrigin()
        self.uniqueNcrs = [min(self.ncrs[a], self.crsInterval[a]) for a in range(3)]

    def _calculateOrigin(self):
        """xyz of voxel (0,0,0) (pdb_eda/ccp4.py:272-286)."""
        if self.futureUse[-3] == 0.0 and self.futureUse[-2] == 0.0 and self.futureUse[-1] == 0.0:
            return np.dot(self.orthoMat, [self.crsStart[self.map2xyz[i]] / self.xyzInterval[i] for i in range(3)])
        # The reference keeps this branch's origin as a Python LIST, so that `origin + [r, r, r]` in getSphereCrsFromXyz
        # (pdb_eda/cutils.pyx:239) concatenates instead of adding and every sphere collapses to a 2-voxel-wide box (and
        # raises in skewed cells): a latent bug that this library does NOT reproduce -- spheres on such
        # maps are enumerated with their real radius.  DESIGN.md section 4 records the deviation.
        import warnings
        warnings.warn("CCP4 header carries a non-zero EM origin (words 50-52): the reference's sphere enumeration degenerates on "
                      "such maps (pdb_eda/cutils.pyx:239); this library uses the real sphere radius", stacklevel=3)
        return [self.originEM[i] for i in range(3)]

    @property
    def orthogonal(self):
        return self.alpha == self.beta == self.gamma == 90

    def xyz2crsCoord(self, xyzCoord):
        """Nearest grid index of an xyz point (pdb_eda/ccp4.py:288-302).  Scalar host arithmetic: this is header
        geometry, not a voxel loop; the batched form is :meth:`DensityMatrix.xyz2crsCoords` on the device."""
        if self.orthogonal:
            grid = [int(round((xyzCoord[i] - self.origin[i]) / self.gridLength[i])) for i in range(3)]
        else:
            frac = np.dot(self.deOrthoMat, xyzCoord)
            grid = [int(round(frac[i] * self.xyzInterval[i])) - self.crsStart[self.map2xyz[i]] for i in range(3)]
        return [grid[self.map2crs[a]] for a in range(3)]

    def crs2xyzCoord(self, crsCoord):
        """xyz of a grid index (pdb_eda/ccp4.py:304-316)."""
        if self.orthogonal:
            return [crsCoord[self.map2xyz[i]] * self.gridLength[i] + self.origin[i] for i in range(3)]
        return np.dot(self.orthoMat,
                      [(crsCoord[self.map2xyz[i]] + self.crsStart[self.map2xyz[i]]) / self.xyzInterval[i] for i in range(3)])


class _TrackedDensity(np.ndarray):
    """float64 view of the voxels that tells its DensityMatrix when it is written to, so the HBM copy is refreshed
    before the next kernel (the reference's tests write into ``density``, tests/test_ccp4.py:77-87)."""

    _owner = None

    def __array_finalize__(self, obj):
        self._owner = getattr(obj, "_owner", None)

    def _touch(self):
        if self._owner is not None:
            self._owner._hostDirty = True

    def __setitem__(self, key, value):
        self._touch()
        super().__setitem__(key, value)

    def __array_ufunc__(self, ufunc, method, *inputs, out=None, **kwargs):
        plain = tuple(np.asarray(x) if isinstance(x, _TrackedDensity) else x for x in inputs)
        if out is not None:
            for o in out:
                if isinstance(o, _TrackedDensity):
                    o._touch()
            kwargs["out"] = tuple(np.asarray(o) if isinstance(o, _TrackedDensity) else o for o in out)
        result = getattr(ufunc, method)(*plain, **kwargs)
        if out is not None and len(out) == 1 and isinstance(out[0], _TrackedDensity):
            return out[0]
        return result


class DensityMatrix:
    """One CCP4 map: header + voxels (pdb_eda/ccp4.py:319-485), voxels resident in HBM."""

    def __init__(self, header, origin, density, pdbid):
        self.pdbid = pdbid
        self.header = header
        self.origin = origin
        n = header.ncrs[0] * header.ncrs[1] * header.ncrs[2]
        flat = np.asarray(density)
        if flat.dtype != np.float32 or flat.dtype.byteorder == ">":
            narrowed = flat.astype(np.float32)
            if flat.dtype.kind == "f" and flat.dtype.itemsize > 4 and not np.array_equal(narrowed.astype(flat.dtype), flat, equal_nan=True):
                raise ValueError("voxel values are not float32-representable; CCP4 mode-2 maps are float32")
            flat = narrowed
        flat = np.ascontiguousarray(flat).reshape(-1)
        if flat.size != n:
            raise ValueError("map payload holds %d values, header says %d" % (flat.size, n))
        self._host32 = flat
        self._density64 = None
        self._hostDirty = False
        self._device = None
        self._totalAbsDensity = {}

    @classmethod
    def _fromDevice(cls, header, origin, rho, pdbid):
        """A map whose voxels are already in HBM (``rho``: flat float32 CUDA tensor, file order) and have no host copy yet."""
        from ._device import DeviceMap, geom_from_header
        self = cls.__new__(cls)
        self.pdbid = pdbid
        self.header = header
        self.origin = origin
        if rho.numel() != header.ncrs[0] * header.ncrs[1] * header.ncrs[2]:
            raise ValueError("map payload holds %d values, header says %d" % (rho.numel(), header.ncrs[0] * header.ncrs[1] * header.ncrs[2]))
        self._host32 = None
        self._density64 = None
        self._hostDirty = False
        self._device = DeviceMap(geom_from_header(header, origin), rho)
        self._totalAbsDensity = {}
        return self

    # ---- host views -----------------------------------------------------------------------------------------
    @property
    def _raw32(self):
        """The flat float32 host copy; read back from HBM on first use when the map was streamed there by ``parse``."""
        if self._host32 is None:
            self._host32 = self._device.rho.cpu().numpy()
        return self._host32

    @_raw32.setter
    def _raw32(self, value):
        self._host32 = value

    @property
    def densityArray(self):
        """Flat voxel values in file order (the reference keeps a tuple, pdb_eda/ccp4.py:337)."""
        self._syncHost()
        return self._raw32

    @property
    def density(self):
        """float64 ``density[section][row][column]`` (pdb_eda/ccp4.py:338); writes are pushed to HBM lazily."""
        if self._density64 is None:
            shape = (self.header.ncrs[2], self.header.ncrs[1], self.header.ncrs[0])
            arr = self._raw32.astype(np.float64).reshape(shape).view(_TrackedDensity)
            arr._owner = self
            self._density64 = arr
        return self._density64

    @density.setter
    def density(self, value):
        shape = (self.header.ncrs[2], self.header.ncrs[1], self.header.ncrs[0])
        arr = np.array(value, dtype=np.float64).reshape(shape).view(_TrackedDensity)
        arr._owner = self
        self._density64 = arr
        self._hostDirty = True

    def _syncHost(self):
        if self._hostDirty:
            self._raw32 = np.ascontiguousarray(np.asarray(self._density64), dtype=np.float32).reshape(-1)
            self._hostDirty = False
            self._device = None
            self._totalAbsDensity = {}

    @property
    def deviceMap(self):
        """The HBM-resident map (:class:`pdb_eda_b200._device.DeviceMap`), uploaded on first use."""
        self._syncHost()
        if self._device is None:
            from ._device import DeviceMap
            self._device = DeviceMap.from_host(self.header, self._raw32, self.origin)
        return self._device

    # ---- statistics -----------------------------------------------------------------------------------------
    @property
    def meanDensity(self):
        """np.mean over all stored voxels (pdb_eda/ccp4.py:343-352), one float64 device reduction."""
        return self.deviceMap.mean_std()[0]

    @property
    def stdDensity(self):
        """np.std (population) over all stored voxels (pdb_eda/ccp4.py:354-363)."""
        return self.deviceMap.mean_std()[1]

    def getTotalAbsDensity(self, densityCutoff):
        """Sum of |rho| over voxels with |rho| > cutoff, cached per cutoff (pdb_eda/ccp4.py:365-376)."""
        self._syncHost()
        if densityCutoff not in self._totalAbsDensity:
            self._totalAbsDensity[densityCutoff] = utils.sumOfAbs(self, densityCutoff)
        return self._totalAbsDensity[densityCutoff]

    # ---- point and sphere queries ---------------------------------------------------------------------------
    def getPointDensityFromCrs(self, crsCoord):
        return utils.getPointDensityFromCrs(self, crsCoord)

    def getPointDensityFromXyz(self, xyzCoord):
        return utils.getPointDensityFromCrs(self, self.header.xyz2crsCoord(xyzCoord))

    def getSphereCrsFromXyz(self, xyzCoord, radius, densityCutoff=0):
        return utils.getSphereCrsFromXyz(self, xyzCoord, radius, densityCutoff)

    def getTotalDensityFromXyz(self, xyzCoord, radius, densityCutoff=0):
        """Total density of a sphere (pdb_eda/ccp4.py:418-435) as one fused enumerate + gather-sum kernel."""
        return utils.getTotalDensityFromXyz(self, xyzCoord, radius, densityCutoff)

    def findAberrantBlobs(self, xyzCoords, radius, densityCutoff=0):
        """Blobs of the in-sphere voxels passing the cutoff (pdb_eda/ccp4.py:437-461)."""
        if not isinstance(xyzCoords[0], (np.floating, float)):
            if len(xyzCoords) > 1:
                crsCoordList = list(utils.getSphereCrsFromXyzList(self, xyzCoords, radius, densityCutoff))
            else:
                crsCoordList = utils.getSphereCrsFromXyz(self, xyzCoords[0], radius, densityCutoff)
        else:
            crsCoordList = utils.getSphereCrsFromXyz(self, xyzCoords, radius, densityCutoff)
        return self.createBlobList(crsCoordList)

    def createFullBlobList(self, cutoff):
        """All blobs of the map beyond ``cutoff`` (pdb_eda/ccp4.py:463-473): fused threshold + CCL + statistics."""
        if cutoff == 0 or float(np.float32(cutoff)) == 0.0:
            return None
        res = utils.fullBlobs(self, cutoff if cutoff > 0 else 0.0, cutoff if cutoff < 0 else 0.0)
        return res[0] if cutoff > 0 else res[1]

    def createFullBlobLists(self, positiveCutoff, negativeCutoff):
        """(green, red) blob lists from ONE pass over the map: createFullBlobList(+c) and createFullBlobList(-c)."""
        return tuple(utils.fullBlobs(self, positiveCutoff, negativeCutoff))

    def createBlobList(self, crsList):
        """Blobs of an arbitrary voxel list (pdb_eda/ccp4.py:475-485)."""
        return utils.blobsFromCrsList(self, crsList)

    # ---- batched forms (no counterpart in the reference; what DensityAnalysis uses) -------------------------------
    def xyz2crsCoords(self, xyzCoords):
        return self.deviceMap.xyz2crs(xyzCoords).cpu().numpy()

    def crs2xyzCoords(self, crsCoords):
        return self.deviceMap.crs2xyz(crsCoords).cpu().numpy()


class DensityBlob:
    """A connected set of voxels with its aggregate properties (pdb_eda/ccp4.py:488-594).

    ``crsList`` is a set of (c, r, s) tuples as in the reference; blobs that come from the device keep their
    voxels as an int32 array and build the set on first access.
    """

    def __init__(self, centroid, coordCenter, totalDensity, volume, crsList, densityMatrix, atoms=None):
        self.centroid = centroid
        self.coordCenter = coordCenter
        self.totalDensity = totalDensity
        self.volume = volume
        self._crsArray = None
        if isinstance(crsList, np.ndarray):
            self._crsArray = crsList
            self._crsSet = None
        else:
            self._crsSet = {tuple(crs) for crs in crsList}
        self.densityMatrix = densityMatrix
        self.atoms = [] if not atoms else atoms

    @property
    def crsList(self):
        if self._crsSet is None:
            self._crsSet = set(map(tuple, self._crsArray.tolist()))
        return self._crsSet

    @crsList.setter
    def crsList(self, value):
        self._crsSet = value
        self._crsArray = None

    @property
    def crsArray(self):
        """The voxels as an (n, 3) int32 array (order unspecified)."""
        if self._crsArray is None or (self._crsSet is not None and len(self._crsSet) != len(self._crsArray)):
            self._crsArray = np.array(sorted(self._crsSet), dtype=np.int32).reshape(-1, 3)
        return self._crsArray

    def __len__(self):
        return len(self._crsArray) if self._crsSet is None else len(self._crsSet)

    @property
    def validCrs(self):
        return utils.testValidCrsList(self.densityMatrix, self.crsList)

    @staticmethod
    def fromCrsList(crsList, densityMatrix):
        """Blob of one voxel list: total density, density-weighted centroid, plain centre, volume
        (pdb_eda/ccp4.py:522-545)."""
        return utils.blobFromCrsList(densityMatrix, crsList)

    def __eq__(self, otherBlob):
        if abs(self.volume - otherBlob.volume) >= 1e-6:
            return False
        if abs(self.totalDensity - otherBlob.totalDensity) >= 1e-6:
            return False
        return all(abs(self.centroid[i] - otherBlob.centroid[i]) < 1e-6 for i in range(3))

    __hash__ = None

    def testOverlap(self, otherBlob):
        return utils.testOverlap(self, otherBlob)

    def merge(self, otherBlob):
        """Union with another blob; aggregates are recomputed over the union (pdb_eda/ccp4.py:575-586)."""
        union = set(self.crsList)
        union.update(otherBlob.crsList)
        atoms = self.atoms + [atom for atom in otherBlob.atoms if atom not in self.atoms]
        merged = DensityBlob.fromCrsList(union, self.densityMatrix)
        self.__dict__.update(merged.__dict__)
        self.atoms = atoms

    def clone(self):
        return DensityBlob(self.centroid, self.coordCenter, self.totalDensity, self.volume,
                           set(self.crsList) if self._crsSet is not None else self._crsArray, self.densityMatrix,
                           self.atoms.copy())
