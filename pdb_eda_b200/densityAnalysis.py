"""Electron-density analysis API -- mirror of ``pdb_eda.densityAnalysis`` with the voxel work on a B200.

Same public surface as the reference (pdb_eda/densityAnalysis.py): ``fromPDBid``, ``fromFile``, ``setGlobals``,
``DensityAnalysis`` with its lazy properties (``symmetryAtoms``, ``greenBlobList``, ``redBlobList``,
``densityElectronRatio``, ...) and methods (``aggregateCloud``, ``calculateAtomSpecificBlobStatistics``, the atom /
residue / symmetry-atom region density and discrepancy calculations incl. the atom-mask path).  What changes is how
the voxel loops run: atoms are batched and every sphere enumeration, gather-sum, clustering, set-union and
nearest-atom search is a CUDA kernel of ``libpdbeda_b200.so``; the Python here only arranges inputs in the
reference's order and formats its result rows.  north_star aliases: ``calcSymmetryAtoms``, ``calcAtomBlobDists``.
"""
import collections
import gzip
import importlib.util
import json
import os
import urllib.request

import numpy as np
import torch

from . import ccp4
from . import cutils as utils
from . import pdbParser
from . import structure as _structure
from . import _device

# ---------------------------------------------------------------------------------------------------- parameters
paramsGlobal = None
radiiGlobal = slopesGlobal = bondedAtomsGlobal = None
fullAtomNameMapElectronsGlobal = fullAtomNameMapAtomTypeGlobal = None
atomTypeLengthGlobal = 0
elementElectronsGlobal = None
masterFullAtomNameMapElectronsGlobal = None


def _referenceConfPath(name):
    """conf/<name> of an installed pdb_eda, if any: the optimised parameter set ships with the reference package
    (pdb_eda/conf/optimized_params.json, loaded at pdb_eda/densityAnalysis.py:32-46), not with this library."""
    override = os.environ.get("PDB_EDA_PARAMS")
    if override and name == "optimized_params.json" and os.path.isfile(override):
        return override
    try:
        spec = importlib.util.find_spec("pdb_eda")
    except (ImportError, ValueError):
        spec = None
    if spec is not None and spec.submodule_search_locations:
        path = os.path.join(list(spec.submodule_search_locations)[0], "conf", name)
        if os.path.isfile(path):
            return path
    return None


def setGlobals(params):
    """Sets the global parameters (radii, slopes, bonded atoms, electrons, atom types) -- pdb_eda/densityAnalysis.py:48-68."""
    global paramsGlobal, radiiGlobal, slopesGlobal, bondedAtomsGlobal, fullAtomNameMapElectronsGlobal
    global fullAtomNameMapAtomTypeGlobal, atomTypeLengthGlobal
    paramsGlobal = params
    radiiGlobal = params["radii"]
    slopesGlobal = params["slopes"]
    bondedAtomsGlobal = params["bonded_atoms"]
    fullAtomNameMapElectronsGlobal = params["full_atom_name_map_electrons"]
    fullAtomNameMapAtomTypeGlobal = params["full_atom_name_map_atom_type"]
    atomTypeLengthGlobal = max(len(atomType) for atomType in fullAtomNameMapAtomTypeGlobal.values()) + 5


def _loadDefaultParams():
    path = _referenceConfPath("optimized_params.json")
    if path is not None:
        with open(path, "r") as fh:
            return json.load(fh)
    from . import synthetic
    return synthetic.defaultParams()  # poly-ALA only; call setGlobals() with a full parameter set for real entries


setGlobals(_loadDefaultParams())


def loadF000Parameters():
    """Loads the F000 electron tables of the reference package (pdb_eda/densityAnalysis.py:70-78)."""
    global elementElectronsGlobal, masterFullAtomNameMapElectronsGlobal
    path = _referenceConfPath("f000_parameters.json.gz")
    if path is None:
        raise RuntimeError("f000_parameters.json.gz ships with the pdb_eda package, which is not installed")
    with gzip.open(path, "rt") as gzipFile:
        f000Params = json.load(gzipFile)
    elementElectronsGlobal = f000Params["element_map_electrons"]
    masterFullAtomNameMapElectronsGlobal = f000Params["full_atom_name_map_electrons"]


ccp4urlPrefix = "http://www.ebi.ac.uk/pdbe/coordinates/files/"
ccp4folder = "./ccp4_data/"
pdbfolder = "./pdb_data/"
pdburlPrefix = "https://files.wwpdb.org/pub/pdb/data/structures/all/pdb/"
mmcifurlPrefix = "http://ftp.rcsb.org/pub/pdb/data/structures/all/mmCIF/"


def _parseStructure(pdbid, handle):
    try:
        import Bio.PDB as biopdb
        if hasattr(biopdb, "PDBParser"):
            return biopdb.PDBParser(QUIET=True).get_structure(pdbid, handle)
    except ImportError:
        pass
    return _structure.parsePDB(handle, pdbid)


def _attachCutoffs(densityObj, diffDensityObj):
    if densityObj is not None:
        densityObj.densityCutoff = densityObj.meanDensity + 1.5 * densityObj.stdDensity
        densityObj.densityCutoffFromHeader = densityObj.header.densityMean + 1.5 * densityObj.header.rmsd
    if diffDensityObj is not None:
        diffDensityObj.diffDensityCutoff = diffDensityObj.meanDensity + 3 * diffDensityObj.stdDensity


def _fetch(url, path):
    folder = os.path.dirname(path)
    if folder and not os.path.exists(folder):
        os.makedirs(folder)
    if not os.path.isfile(path):
        urllib.request.urlretrieve(url, path)
    return path


def fromPDBid(pdbid, ccp4density=True, ccp4diff=True, pdbbio=True, pdbi=True, downloadFile=True, mmcif=False):
    """DensityAnalysis of a PDB entry, downloading (and caching) its maps and coordinates; 0 on any failure
    (pdb_eda/densityAnalysis.py:88-179)."""
    pdbid = pdbid.lower()
    densityObj = diffDensityObj = pdbObj = biopdbObj = None
    try:
        if ccp4density:
            if downloadFile:
                densityObj = ccp4.read(_fetch(ccp4urlPrefix + pdbid + ".ccp4", ccp4folder + pdbid + ".ccp4"), pdbid)
            else:
                densityObj = ccp4.readFromPDBID(pdbid)
        if ccp4diff:
            if downloadFile:
                diffDensityObj = ccp4.read(_fetch(ccp4urlPrefix + pdbid + "_diff.ccp4", ccp4folder + pdbid + "_diff.ccp4"), pdbid)
            else:
                diffDensityObj = ccp4.readFromPDBID(pdbid + "_diff")
        _attachCutoffs(densityObj, diffDensityObj)
        if pdbbio or pdbi:
            pdbfile = _fetch(pdburlPrefix + "pdb" + pdbid + ".ent.gz", pdbfolder + "pdb" + pdbid + ".ent.gz")
            if pdbbio:
                with gzip.open(pdbfile, "rt") as gzipFile:
                    biopdbObj = _parseStructure(pdbid, gzipFile)
            if pdbi:
                with gzip.open(pdbfile, "rt") as gzipFile:
                    pdbObj = pdbParser.readPDBfile(gzipFile)
        if mmcif and downloadFile:
            _fetch(mmcifurlPrefix + pdbid + ".cif.gz", pdbfolder + pdbid + ".cif.gz")
    except Exception:
        return 0
    return DensityAnalysis(pdbid, densityObj, diffDensityObj, biopdbObj, pdbObj)


def fromFile(pdbFile, ccp4DensityFile=None, ccp4DiffDensityFile=None):
    """DensityAnalysis from local files or open handles; 0 on any failure (pdb_eda/densityAnalysis.py:182-229)."""
    pdbid = "xxxx"
    densityObj = diffDensityObj = None
    try:
        if ccp4DensityFile is not None:
            densityObj = ccp4.read(ccp4DensityFile, pdbid) if isinstance(ccp4DensityFile, str) else ccp4.parse(ccp4DensityFile, pdbid)
        if ccp4DiffDensityFile is not None:
            diffDensityObj = (ccp4.read(ccp4DiffDensityFile, pdbid) if isinstance(ccp4DiffDensityFile, str)
                              else ccp4.parse(ccp4DiffDensityFile, pdbid))
        _attachCutoffs(densityObj, diffDensityObj)
        if isinstance(pdbFile, str) and pdbFile.endswith(".gz"):
            with gzip.open(pdbFile, "rt") as gzipFile:
                biopdbObj = _parseStructure(pdbid, gzipFile)
            with gzip.open(pdbFile, "rt") as gzipFile:
                pdbObj = pdbParser.readPDBfile(gzipFile)
        elif isinstance(pdbFile, str):
            with open(pdbFile, "r") as handle:
                biopdbObj = _parseStructure(pdbid, handle)
            pdbObj = pdbParser.readPDBfile(pdbFile)
        else:
            text = pdbFile.read()
            import io
            biopdbObj = _parseStructure(pdbid, io.StringIO(text))
            pdbObj = pdbParser.readPDBfile(io.StringIO(text))
    except Exception:
        return 0
    return DensityAnalysis(pdbid, densityObj, diffDensityObj, biopdbObj, pdbObj)


def cleanPDBid(pdbid):
    """Removes the cached files of an entry (pdb_eda/densityAnalysis.py:232-259)."""
    pdbid = pdbid.lower()
    try:
        for path in (ccp4folder + pdbid + ".ccp4", ccp4folder + pdbid + "_diff.ccp4", pdbfolder + "pdb" + pdbid + ".ent.gz",
                     pdbfolder + pdbid + ".cif.gz"):
            if os.path.isfile(path):
                os.remove(path)
    except Exception:
        return False
    return True


def testCCP4URL(pdbid):
    """Whether the PDBe API has electron density statistics for the entry (pdb_eda/densityAnalysis.py:262-275)."""
    try:
        urllib.request.urlopen("https://www.ebi.ac.uk/pdbe/api/pdb/entry/electron_density_statistics/" + pdbid)
    except urllib.request.HTTPError:
        return False
    return True


def residueAtomName(atom):
    """RESNAME_ATOMNAME key into the parameter tables (pdb_eda/densityAnalysis.py:1243-1251)."""
    return atom.parent.resname.strip() + "_" + atom.name


def _median(t):
    """np.median of a 1-D device tensor (mean of the two middle values for an even count)."""
    n = t.numel()
    if n == 0:
        return float("nan")
    s = torch.sort(t).values
    return float(s[n // 2].item()) if n % 2 else float((s[n // 2 - 1].item() + s[n // 2].item()) / 2.0)


class _FcMatrix(ccp4.DensityMatrix):
    """The reference's ``fc`` object: a deep copy of the 2Fo-Fc DensityMatrix whose ``density`` is replaced
    (pdb_eda/densityAnalysis.py:426-435) -- header, ``densityArray``, mean and standard deviation stay those of 2Fo-Fc."""

    def __init__(self, densityObj, diffDensityObj):
        self.pdbid = densityObj.pdbid
        self.header = densityObj.header
        self.origin = densityObj.origin
        self._fo = densityObj
        self._diff = diffDensityObj
        self._density64 = None
        self._hostDirty = False
        self._totalAbsDensity = {}

    @property
    def densityArray(self):
        return self._fo.densityArray

    @property
    def density(self):
        if self._density64 is None:
            self._density64 = np.asarray(self._fo.density) - np.asarray(self._diff.density) * 2
        return self._density64

    @property
    def meanDensity(self):
        return self._fo.meanDensity

    @property
    def stdDensity(self):
        return self._fo.stdDensity

    @property
    def deviceMap(self):
        raise NotImplementedError("the Fc map is formed on the fly from the 2Fo-Fc and Fo-Fc maps; use the RSCC / RSR methods")


def _components(n, neighbours):
    """Connected components of an overlap graph, grown exactly like the reference grows them (Python sets, the same
    insertion sequence, pdb_eda/densityAnalysis.py:664-673 and :696-705), so that set iteration order -- which decides
    the base cloud and the atom order of a merged cloud -- is the reference's."""
    used = set()
    comps = []
    for start in range(n):
        if start not in used:
            newCluster = {index for index in neighbours[start]}
            currCluster = set([start])
            currCluster.update(newCluster)
            while len(newCluster):
                newCluster = {index for oldIndex in newCluster for index in neighbours[oldIndex] if index not in currCluster}
                currCluster.update(newCluster)
            used.update(currCluster)
            comps.append(currCluster)
    return comps


def _neighbourLists(n, pairs):
    nbrs = [[] for _ in range(n)]
    for a, b in pairs:
        nbrs[a].append(b)
        nbrs[b].append(a)
    for lst in nbrs:
        lst.sort()
    return nbrs


class DensityAnalysis(object):
    """Density, difference density, structure and PDB header of one entry plus the analyses on them
    (pdb_eda/densityAnalysis.py:278-1241)."""

    def __init__(self, pdbid, densityObj=None, diffDensityObj=None, biopdbObj=None, pdbObj=None):
        self.pdbid = pdbid
        self.densityObj = densityObj
        self.diffDensityObj = diffDensityObj
        self.biopdbObj = biopdbObj
        self.pdbObj = pdbObj
        for name in ("symmetryAtoms", "symmetryOnlyAtoms", "asymmetryAtoms", "symmetryAtomCoords", "symmetryOnlyAtomCoords",
                     "asymmetryAtomCoords", "greenBlobList", "redBlobList", "blueBlobList", "fc", "medians",
                     "atomCloudDescriptions", "residueCloudDescriptions", "domainCloudDescriptions", "F000",
                     "densityElectronRatio", "numVoxelsAggregated", "totalAggregatedElectrons", "totalAggregatedDensity",
                     "atomTypeOverlapCompleteness", "atomTypeOverlapIncompleteness"):
            setattr(self, "_" + name, None)

    def resetCloud(self):
        """Forgets the cloud-aggregation results (after ``setGlobals`` changed the radii); maps stay resident in HBM."""
        for name in ("medians", "atomCloudDescriptions", "residueCloudDescriptions", "domainCloudDescriptions", "densityElectronRatio",
                     "numVoxelsAggregated", "totalAggregatedElectrons", "totalAggregatedDensity", "atomTypeOverlapCompleteness",
                     "atomTypeOverlapIncompleteness"):
            setattr(self, "_" + name, None)

    # ------------------------------------------------------------------------------------------ lazy properties
    def _symmetry(self, name):
        if self._symmetryAtoms is None:
            self._calculateSymmetryAtoms()
        return getattr(self, name)

    symmetryAtoms = property(lambda self: self._symmetry("_symmetryAtoms"))
    symmetryOnlyAtoms = property(lambda self: self._symmetry("_symmetryOnlyAtoms"))
    asymmetryAtoms = property(lambda self: self._symmetry("_asymmetryAtoms"))
    symmetryAtomCoords = property(lambda self: self._symmetry("_symmetryAtomCoords"))
    symmetryOnlyAtomCoords = property(lambda self: self._symmetry("_symmetryOnlyAtomCoords"))
    asymmetryAtomCoords = property(lambda self: self._symmetry("_asymmetryAtomCoords"))

    def _greenRed(self):
        """Both difference-map blob lists from one pass over the map (pdb_eda/densityAnalysis.py:392-412)."""
        cut = self.diffDensityObj.diffDensityCutoff
        self._greenBlobList, self._redBlobList = self.diffDensityObj.createFullBlobLists(cut, -1 * cut)

    @property
    def greenBlobList(self):
        if self._greenBlobList is None:
            self._greenRed()
        return self._greenBlobList

    @property
    def redBlobList(self):
        if self._redBlobList is None:
            self._greenRed()
        return self._redBlobList

    @property
    def blueBlobList(self):
        if self._blueBlobList is None:
            self._blueBlobList = self.densityObj.createFullBlobList(self.densityObj.densityCutoff)
        return self._blueBlobList

    @property
    def fc(self):
        """Fc map = 2Fo-Fc - 2 (Fo-Fc) (pdb_eda/densityAnalysis.py:426-435).  Like the reference's deep copy it keeps the
        header, mean and standard deviation of the 2Fo-Fc object; only ``density`` (float64, not float32-representable)
        differs.  The device kernels form it on the fly from the two resident maps (``pe_pair_metrics``)."""
        if self._fc is None:
            self._fc = _FcMatrix(self.densityObj, self.diffDensityObj)
        return self._fc

    @property
    def fo(self):
        return self.densityObj

    @property
    def F000(self):
        if self._F000 is None:
            self._F000 = self.estimateF000()
        return self._F000

    def _cloud(self, name):
        if getattr(self, name) is None:
            self.aggregateCloud()
        return getattr(self, name)

    medians = property(lambda self: self._cloud("_medians"))
    atomCloudDescriptions = property(lambda self: self._cloud("_atomCloudDescriptions"))
    residueCloudDescriptions = property(lambda self: self._cloud("_residueCloudDescriptions"))
    domainCloudDescriptions = property(lambda self: self._cloud("_domainCloudDescriptions"))
    numVoxelsAggregated = property(lambda self: self._cloud("_numVoxelsAggregated"))
    totalAggregatedElectrons = property(lambda self: self._cloud("_totalAggregatedElectrons"))
    totalAggregatedDensity = property(lambda self: self._cloud("_totalAggregatedDensity"))
    densityElectronRatio = property(lambda self: self._cloud("_densityElectronRatio"))
    atomTypeOverlapCompleteness = property(lambda self: self._cloud("_atomTypeOverlapCompleteness"))
    atomTypeOverlapIncompleteness = property(lambda self: self._cloud("_atomTypeOverlapIncompleteness"))

    residueCloudHeader = ['chain', 'residue_number', 'residue_name', 'local_density_electron_ratio', 'num_voxels', 'electrons',
                          'volume', 'centroid_xyz']
    domainCloudHeader = residueCloudHeader

    # ------------------------------------------------------------------------------------------ cloud aggregation
    def aggregateCloud(self, minCloudElectrons=25.0, minTotalElectrons=400.0):
        """Aggregates the 2Fo-Fc map into atom, residue and domain clouds and estimates the density-electron ratio
        (pdb_eda/densityAnalysis.py:571-780).

        Device pipeline: one batched sphere enumeration with per-atom 26-connected clustering gives every atom's
        clouds; the pairwise cloud overlaps of a residue and of the whole structure come from one hash-table pass
        each (instead of O(|A| |B|) generator loops per pair); merged residue / domain clouds are voxel SETS, i.e.
        connected components of the union, whose sums are taken over each distinct voxel once.
        """
        densityObj = self.densityObj
        types, electronsOf = fullAtomNameMapAtomTypeGlobal, fullAtomNameMapElectronsGlobal
        unitVolume = densityObj.header.unitVolume
        # ---- candidate atoms in the reference's traversal order (:596-603)
        cand = []        # (residue index, atom, resAtom)
        residues = []
        for residue in self.biopdbObj.get_residues():
            if residue.id[0] != ' ':
                continue
            ridx = len(residues)
            residues.append(residue)
            for atom in residue.child_list:
                resAtom = residueAtomName(atom)
                if resAtom not in types or atom.get_occupancy() == 0:
                    continue
                cand.append((ridx, atom, resAtom))
        completelyOverlappedAtomTypes = collections.defaultdict(int)
        incompletelyOverlappedAtomTypes = collections.defaultdict(int)
        nAtoms = len(cand)
        coords32 = np.array([c[1].coord for c in cand], dtype=np.float32).reshape(-1, 3)
        coords = coords32.astype(np.float64)
        radii = np.array([radiiGlobal[types[c[2]]] for c in cand], dtype=np.float64)

        # ---- pass 1: every atom's clouds = findAberrantBlobs(atom.coord, radius, densityCutoff) (:605)
        lists = utils.sphereLists(densityObj, coords, radii.astype(np.float32) if nAtoms else radii, densityObj.densityCutoff,
                                  values=False, labels=True)
        vAtom = lists["atom"].long()
        vLabel = lists["label"].long()
        vCrs = lists["crs"]
        dev = vCrs.device
        nClouds = torch.zeros(nAtoms, dtype=torch.int64, device=dev)
        if len(vAtom):
            nClouds.scatter_reduce_(0, vAtom, vLabel + 1, reduce="amax")
        cloudStart = torch.zeros(nAtoms + 1, dtype=torch.int64, device=dev)
        cloudStart[1:] = torch.cumsum(nClouds, 0)
        vCloud = cloudStart[vAtom] + vLabel                      # global cloud id (atom order, creation order inside)
        totalClouds = int(cloudStart[-1].item())
        cstats = utils.crsStats(densityObj, vCrs, vCloud.to(torch.int32), None, totalClouds).cpu().numpy()
        cloudStartH = cloudStart.cpu().numpy()
        nCloudsH = np.diff(cloudStartH)
        with np.errstate(divide="ignore", invalid="ignore"):
            cCentroid = cstats[:, 2:5] / cstats[:, 1:2]
        cTotal, cCount = cstats[:, 1], cstats[:, 0].astype(np.int64)
        cloudAtom = np.repeat(np.arange(nAtoms), nCloudsH)
        delta = coords[cloudAtom] - cCentroid
        cDist = np.sqrt((delta * delta).sum(axis=1))              # np.linalg.norm(atom.coord - cloud.centroid)
        # allAtomClouds is keyed by the coordinate tuple (:606): a later atom with identical coordinates replaces the entry
        if nAtoms:
            _, inverse = np.unique(coords32 + np.float32(0.0), axis=0, return_inverse=True)   # -0.0 and 0.0 are one key
            inverse = np.asarray(inverse).reshape(-1)
            lastOfGroup = np.zeros(int(inverse.max()) + 1, dtype=np.int64)
            lastOfGroup[inverse] = np.arange(nAtoms)             # later assignments win: the last atom with these coordinates
            src = lastOfGroup[inverse]
        else:
            src = np.zeros(0, np.int64)
        centroidDistances = np.minimum.reduceat(cDist, cloudStartH[:-1][nCloudsH > 0]) if totalClouds else []
        if len(centroidDistances):
            centroidDistanceCutoff = np.nanmedian(centroidDistances) + 2.5 * np.nanstd(centroidDistances)
        else:
            centroidDistanceCutoff = np.nan

        # ---- pass 2, host part: which atoms contribute, their best cloud, the residue pools (:611-643), vectorised
        # per candidate: the clouds it uses are those stored under its coordinates (those of atom src[k])
        nC = nCloudsH[src] if nAtoms else np.zeros(0, np.int64)
        lo = cloudStartH[:-1][src] if nAtoms else np.zeros(0, np.int64)
        # first minimum of the centroid distances per cloud owner (distances.index(min(distances)), :630-634)
        minD = np.full(nAtoms, np.inf)
        firstArg = np.zeros(nAtoms, dtype=np.int64)
        if totalClouds:
            owners = np.flatnonzero(nCloudsH > 0)
            minD[owners] = centroidDistances
            isMin = cDist == minD[cloudAtom]
            firstArg[owners] = np.minimum.reduceat(np.where(isMin, np.arange(totalClouds), totalClouds), cloudStartH[:-1][owners])
        accepted = (nC > 0) & ((nC == 1) | ~(minD[src] > centroidDistanceCutoff)) if nAtoms else np.zeros(0, bool)
        best = np.where(nC == 1, lo, firstArg[src]) if nAtoms else lo
        accIdx = np.flatnonzero(accepted)
        candResidue = np.array([c[0] for c in cand], dtype=np.int64) if nAtoms else np.zeros(0, np.int64)
        # pool entries: every cloud of every accepted atom, in traversal order
        accN = nC[accIdx]
        poolAtomA = np.repeat(accIdx, accN)
        poolStartOfAtom = np.cumsum(accN) - accN
        poolCloudA = np.repeat(lo[accIdx], accN) + (np.arange(int(accN.sum())) - np.repeat(poolStartOfAtom, accN))
        poolResidueA = candResidue[poolAtomA]
        nPool = len(poolCloudA)
        poolAtom = poolAtomA.tolist()
        poolResidue = poolResidueA.tolist()
        resPoolStart = np.searchsorted(poolResidueA, np.arange(len(residues) + 1)).tolist()
        # atom table rows
        bestAcc = best[accIdx]
        rowDensity = cTotal[bestAcc].tolist()
        rowCount = cCount[bestAcc].tolist()
        rowDist = cDist[bestAcc].tolist()
        rowCentroid = cCentroid[bestAcc].tolist()
        atomList = []
        resAtomClouds = [dict() for _ in residues]      # per residue: {resAtom: [local pool indices]}
        for j, k in enumerate(accIdx.tolist()):
            ridx, atom, resAtom = cand[k]
            residue = residues[ridx]
            first = int(poolStartOfAtom[j]) - resPoolStart[ridx]
            resAtomClouds[ridx][resAtom] = list(range(first, first + int(accN[j])))
            electrons = electronsOf[resAtom]
            atomList.append([residue.parent.id, residue.id[1], atom.parent.resname, atom.name, types[resAtom],
                             rowDensity[j] / electrons / atom.get_occupancy(), rowCount[j], electrons, atom.get_bfactor(),
                             rowDist[j], rowCentroid[j]])

        # ---- the voxels of the pool, in pool order (a cloud shared by two atoms with equal coordinates appears twice)
        order = torch.argsort(vCloud, stable=True)
        cloudVoxStart = torch.zeros(totalClouds + 1, dtype=torch.int64, device=dev)
        cloudVoxStart[1:] = torch.cumsum(torch.bincount(vCloud, minlength=totalClouds), 0)
        if nPool:
            pc = torch.from_numpy(poolCloudA).to(dev)
            lens = cloudVoxStart[pc + 1] - cloudVoxStart[pc]
            pOwner = torch.repeat_interleave(torch.arange(nPool, device=dev), lens)
            offs = torch.arange(int(lens.sum().item()), device=dev) - torch.repeat_interleave(torch.cumsum(lens, 0) - lens, lens)
            pIdx = order[cloudVoxStart[pc][pOwner] + offs]
            pCrs = vCrs[pIdx].contiguous()
            pRes = torch.from_numpy(poolResidueA).to(dev)[pOwner]
        else:
            pOwner = torch.zeros(0, dtype=torch.int64, device=dev)
            pCrs = torch.zeros((0, 3), dtype=torch.int32, device=dev)
            pRes = pOwner

        # ---- residue level: pairwise overlaps (:646-649), completeness (:653-659), merged residue clouds (:662-677)
        pairs = utils.overlapPairs(pCrs, pOwner.to(torch.int32), pRes.to(torch.int32)) if nPool else np.zeros((0, 2), np.int32)
        rLabel, rFirst, nResClouds = utils.clusterCrs(pCrs, pRes.to(torch.int32), wantFirst=True) if nPool else (None, None, 0)
        rStats = (utils.crsStats(densityObj, pCrs, rLabel, rFirst, nResClouds).cpu().numpy() if nPool else np.zeros((0, 8)))
        poolFirstVoxel = (torch.cumsum(lens, 0) - lens).cpu().numpy() if nPool else np.zeros(0, np.int64)
        poolResLabel = rLabel.cpu().numpy()[poolFirstVoxel] if nPool else np.zeros(0, np.int64)
        pairsByResidue = collections.defaultdict(list)
        for a, b in pairs.tolist():
            pairsByResidue[poolResidue[a]].append((a, b))
        residueList = []
        domainAtoms = []        # per residue cloud (domain pool order): atoms in the reference's merge order
        domainResLabel = []     # device label of that residue cloud
        electronCache = [electronsOf[c[2]] * c[1].get_occupancy() for c in cand]
        for ridx, residue in enumerate(residues):
            p0, p1 = resPoolStart[ridx], resPoolStart[ridx + 1]
            npool = p1 - p0
            local = [(a - p0, b - p0) for a, b in pairsByResidue.get(ridx, ())]
            overlap = set(local)
            overlap.update((b, a) for a, b in local)
            indices = resAtomClouds[ridx]
            for atom in residue.child_list:
                resAtom = residueAtomName(atom)
                if resAtom in indices:
                    if all(any((i1, i2) in overlap for i1 in indices[resAtom] for i2 in indices[resAtom2])
                           for resAtom2 in bondedAtomsGlobal[resAtom] if resAtom2 in indices):
                        completelyOverlappedAtomTypes[types[resAtom]] += 1
                    else:
                        incompletelyOverlappedAtomTypes[types[resAtom]] += 1
            for currCluster in _components(npool, _neighbourLists(npool, local)):
                base = currCluster.pop()
                atoms = [poolAtom[p0 + base]]
                for idx in currCluster:
                    k = poolAtom[p0 + idx]
                    if k not in atoms:
                        atoms.append(k)
                lab = int(poolResLabel[p0 + base])
                st = rStats[lab]
                resElectrons = sum([electronCache[k] for k in atoms])
                nvox = int(st[0])
                if resElectrons >= minCloudElectrons:
                    residueList.append([residue.parent.id, residue.id[1], residue.resname, st[1] / resElectrons, nvox, resElectrons,
                                        nvox * unitVolume, (st[2:5] / st[1]).tolist()])
                domainAtoms.append(atoms)
                domainResLabel.append(lab)

        # ---- domain level (:689-708): overlaps between residue clouds, merged domain clouds
        nDomPool = len(domainAtoms)
        if nPool:
            labelToDom = np.full(nResClouds, -1, dtype=np.int64)
            labelToDom[np.asarray(domainResLabel, dtype=np.int64)] = np.arange(nDomPool)
            vDom = torch.from_numpy(labelToDom).to(dev)[rLabel.long()]
            dpairs = utils.overlapPairs(pCrs, vDom.to(torch.int32))
            dLabel, dFirst, nDomains = utils.clusterCrs(pCrs, None, wantFirst=True)
            dStats = utils.crsStats(densityObj, pCrs, dLabel, dFirst, nDomains).cpu().numpy()
            dLabelH = dLabel.cpu().numpy()
            domFirstVoxel = np.full(nDomPool, -1, dtype=np.int64)
            vDomH = vDom.cpu().numpy()
            firstIdx = np.unique(vDomH, return_index=True)
            domFirstVoxel[firstIdx[0]] = firstIdx[1]
        else:
            dpairs, dStats = np.zeros((0, 2), np.int32), np.zeros((0, 8))
        domainList = []
        numVoxels = 0
        totalElectrons = 0
        totalDensity = 0
        for currCluster in _components(nDomPool, _neighbourLists(nDomPool, dpairs.tolist())):
            base = currCluster.pop()
            atoms = list(domainAtoms[base])
            seen = set(atoms)
            for idx in currCluster:                      # atoms + [atom for atom in other.atoms if atom not in atoms] (ccp4.py:582)
                for k in domainAtoms[idx]:
                    if k not in seen:
                        seen.add(k)
                        atoms.append(k)
            st = dStats[int(dLabelH[domFirstVoxel[base]])]
            atom = cand[atoms[0]][1]
            domainElectrons = sum([electronCache[k] for k in atoms])
            nvox = int(st[0])
            totalElectrons += domainElectrons
            numVoxels += nvox
            totalDensity += st[1]
            if domainElectrons >= minCloudElectrons:
                domainList.append([atom.parent.parent.id, atom.parent.id[1], atom.parent.resname, st[1] / domainElectrons, nvox,
                                   domainElectrons, nvox * unitVolume, (st[2:5] / st[1]).tolist()])
        if totalElectrons < minTotalElectrons:
            return
        densityElectronRatio = totalDensity / totalElectrons
        domainList.sort(key=lambda x: x[3])

        # ---- per-atom-type statistics (:734-769); host numpy / scipy, as in the reference
        try:
            atoms, medians = self._atomTypeStatistics(atomList, densityElectronRatio, unitVolume)
        except Exception:
            return
        self._densityElectronRatio = densityElectronRatio
        self._numVoxelsAggregated = numVoxels
        self._totalAggregatedElectrons = totalElectrons
        self._totalAggregatedDensity = totalDensity
        self._medians = medians
        self._atomCloudDescriptions = atoms
        self._residueCloudDescriptions = residueList
        self._domainCloudDescriptions = domainList
        self._atomTypeOverlapCompleteness = completelyOverlappedAtomTypes
        self._atomTypeOverlapIncompleteness = incompletelyOverlappedAtomTypes

    @staticmethod
    def _atomTypeStatistics(atomList, densityElectronRatio, unitVolume):
        """The structured atom table, per-type medians and b-factor slopes (pdb_eda/densityAnalysis.py:734-766)."""
        from scipy import stats
        currentSlopes = slopesGlobal
        dataType = np.dtype([('chain', np.dtype(('U', 20))), ('residue_number', int), ('residue_name', np.dtype(('U', 10))),
                             ('atom_name', np.dtype(('U', 10))), ('atom_type', np.dtype(('U', atomTypeLengthGlobal))),
                             ('density_electron_ratio', float), ('num_voxels', int), ('electrons', int), ('bfactor', float),
                             ('centroid_distance', float), ('centroid_xyz', float, (3,)), ('adj_density_electron_ratio', float),
                             ('domain_fraction', float), ('corrected_fraction', float), ('corrected_density_electron_ratio', float),
                             ('volume', float)])
        atoms = np.asarray([tuple(row + [0.0] * 5) for row in atomList], dataType)
        if not np.isnan(atoms['centroid_distance']).all():
            centroidCutoff = np.nanmedian(atoms['centroid_distance']) + np.nanstd(atoms['centroid_distance']) * 2
            atoms = atoms[atoms['centroid_distance'] < centroidCutoff]
        atom_types, typeIndex = np.unique(atoms['atom_type'], return_inverse=True)
        typeIndex = np.asarray(typeIndex).reshape(-1)

        masks = [typeIndex == i for i in range(len(atom_types))]      # atoms['atom_type'] == t, once per type

        def typeMedians(columns, table):
            return {column: {t: np.nanmedian(table[column][m]) for t, m in zip(atom_types, masks)} for column in columns}

        medians = typeMedians(['num_voxels'], atoms)

        def lookup(column, _types):
            """medians[column][atom type] for every row (the reference uses np.vectorize over a dict lookup)."""
            return np.array([medians[column][t] for t in atom_types])[typeIndex]
        atoms['adj_density_electron_ratio'] = atoms['density_electron_ratio'] / atoms['num_voxels'] * lookup('num_voxels', atoms['atom_type'])
        atoms['volume'] = atoms['num_voxels'] * unitVolume
        medians.update(typeMedians(['density_electron_ratio', 'centroid_distance', 'adj_density_electron_ratio', 'volume'], atoms))
        medians['bfactor'] = {t: np.nanmedian(atoms['bfactor'][m & (atoms['bfactor'] > 0)]) for t, m in zip(atom_types, masks)}
        atoms['bfactor'][atoms['bfactor'] <= 0] = lookup('bfactor', atoms['atom_type'])[atoms['bfactor'] <= 0]

        def calcSlope(data, atom_type):
            if len(data['chain']) <= 2 or len(np.unique(data['bfactor'])) == 1:
                return currentSlopes[atom_type]
            fit = stats.linregress(np.log(data['bfactor']), (data['adj_density_electron_ratio'] - densityElectronRatio) / densityElectronRatio)
            return currentSlopes[atom_type] if fit[3] > 0.05 else fit[0]

        medians['slopes'] = {t: calcSlope(atoms[m], t) for t, m in zip(atom_types, masks)}
        atoms['domain_fraction'] = (atoms['adj_density_electron_ratio'] - densityElectronRatio) / densityElectronRatio
        atoms['corrected_fraction'] = atoms['domain_fraction'] - (np.log(atoms['bfactor']) - np.log(lookup('bfactor', atoms['atom_type']))) * lookup('slopes', atoms['atom_type'])
        atoms['corrected_density_electron_ratio'] = atoms['corrected_fraction'] * densityElectronRatio + densityElectronRatio
        medians.update(typeMedians(['domain_fraction', 'corrected_fraction', 'corrected_density_electron_ratio'], atoms))
        return atoms, medians

    # ------------------------------------------------------------------------------------------ RSCC / RSR
    def medianAbsFoFc(self):
        """Median |Fo| and |Fc| over the voxels of the unique volume where both are below one sigma
        (pdb_eda/densityAnalysis.py:783-800); masked selection and medians run on the device."""
        fo, diff = self.densityObj.deviceMap, self.diffDensityObj.deviceMap
        cut = self.densityObj.meanDensity + 1.0 * self.densityObj.stdDensity   # the Fc copy keeps the Fo statistics
        g = fo.geom
        ns, nr, nc = g.ncrs[2], g.ncrs[1], g.ncrs[0]
        u = g.unique_ncrs
        a = fo.rho.view(ns, nr, nc)[:u[2], :u[1], :u[0]].double()
        b = a - diff.rho.view(ns, nr, nc)[:u[2], :u[1], :u[0]].double() * 2
        keep = (a.abs() < cut) & (b.abs() < cut)
        return (_median(a[keep].abs()), _median(b[keep].abs()))

    residueMetricsHeaderList = ['chain', 'residue_number', 'residue_name', "rscc", "rsr", "mean_occupancy", "occupancy_weighted_mean_bfactor"]
    atomMetricsHeaderList = ['chain', 'residue_number', 'residue_name', "atom_name", "symmetry", "xyz", "rscc", "rsr", "occupancy", "bfactor"]

    def _metricsRadius(self):
        resolution = self.biopdbObj.header['resolution']
        radius = 0.7
        if 0.6 <= resolution <= 3:
            radius = (resolution - 0.6) / 3 + 0.7
        elif resolution > 3:
            radius = resolution * 0.5
        return radius

    def _groupMetrics(self, coords, groupOfAtom, nGroups, radius):
        """RSCC / RSR of the set-union of the atoms' spheres per group: one sphere enumeration, one de-duplication and
        one two-pass reduction over both maps for all groups."""
        lists = utils.sphereLists(self.densityObj, coords, radius, 0.0)
        crs = lists["crs"]
        group = torch.from_numpy(np.asarray(groupOfAtom, dtype=np.int32)).to(crs.device)[lists["atom"].long()]
        _, first, _ = utils.clusterCrs(crs, group, wantFirst=True)
        out = utils.pairMetrics(self.densityObj, self.diffDensityObj, crs, group, first, nGroups).cpu().numpy()
        with np.errstate(divide="ignore", invalid="ignore"):
            rscc = np.clip(out[:, 7] / np.sqrt(out[:, 5] * out[:, 6]), -1.0, 1.0)
            rsr = out[:, 3] / out[:, 4]
        return rscc, rsr

    def residueMetrics(self, residueList=None):
        """RSCC and RSR of every residue on the Fo / Fc maps (pdb_eda/densityAnalysis.py:803-829)."""
        radius = self._metricsRadius()
        if residueList is None:
            residueList = list(self.biopdbObj.get_residues())
        coords, groupOf = [], []
        for k, residue in enumerate(residueList):
            for atom in residue.child_list:
                coords.append(atom.coord)
                groupOf.append(k)
        if not coords:
            return []
        rscc, rsr = self._groupMetrics(coords, groupOf, len(residueList), radius)
        results = []
        for k, residue in enumerate(residueList):
            bfactorWeightedSum = occupancySum = 0.0
            for atom in residue.child_list:
                bfactorWeightedSum += atom.get_bfactor() * atom.get_occupancy()
                occupancySum += atom.get_occupancy()
            results.append([residue.parent.id, residue.id[1], residue.resname, rscc[k], rsr[k], occupancySum / len(residue.child_list),
                            bfactorWeightedSum / occupancySum])
        return results

    def atomMetrics(self, atomList=None):
        """RSCC and RSR of every atom (pdb_eda/densityAnalysis.py:831-858)."""
        radius = self._metricsRadius()
        if atomList is None:
            atomList = self.asymmetryAtoms
        if not atomList:
            return []
        rscc, rsr = self._groupMetrics([np.asarray(atom.coord, dtype=np.float64) for atom in atomList], list(range(len(atomList))),
                                       len(atomList), radius)
        return [[atom.parent.parent.id, atom.parent.id[1], atom.parent.resname, atom.name, atom.symmetry, atom.coord, rscc[k], rsr[k],
                 atom.get_occupancy(), atom.get_bfactor()] for k, atom in enumerate(atomList)]

    def calculateRsccRsrMetrics(self, crsList):
        """(RSCC, RSR) over one voxel list (pdb_eda/densityAnalysis.py:860-882)."""
        crs = np.asarray(list(crsList), dtype=np.int32).reshape(-1, 3)
        out = utils.pairMetrics(self.densityObj, self.diffDensityObj, crs, None, None, 1).cpu().numpy()[0]
        with np.errstate(divide="ignore", invalid="ignore"):
            return (float(np.clip(out[7] / np.sqrt(out[5] * out[6]), -1.0, 1.0)), float(out[3] / out[4]))

    # ------------------------------------------------------------------------------------------ symmetry atoms
    def _calculateSymmetryAtoms(self):
        """Symmetry-operator x lattice-translation images of every atom inside the map's circumscribed box +- 5 A
        (pdb_eda/densityAnalysis.py:885-912)."""
        header = self.densityObj.header
        ncrs = header.ncrs
        corners = [header.crs2xyzCoord([c, r, s]) for c in [0, ncrs[0] - 1] for r in [0, ncrs[1] - 1] for s in [0, ncrs[2] - 1]]
        xs = sorted([i[0] for i in corners])
        ys = sorted([i[1] for i in corners])
        zs = sorted([i[2] for i in corners])
        allAtoms = utils.createSymmetryAtoms(list(self.biopdbObj.get_atoms()), self.pdbObj.header.rotationMats, header.orthoMat, xs, ys, zs)
        self._symmetryAtoms = allAtoms
        self._symmetryAtomCoords = np.asarray([atom.coord for atom in allAtoms])
        self._symmetryOnlyAtoms = [atom for atom in allAtoms if atom.symmetry != (0, 0, 0, 0)]
        self._symmetryOnlyAtomCoords = np.asarray([atom.coord for atom in self._symmetryOnlyAtoms])
        self._asymmetryAtoms = [atom for atom in allAtoms if atom.symmetry == (0, 0, 0, 0)]
        self._asymmetryAtomCoords = np.asarray([atom.coord for atom in self._asymmetryAtoms])

    calcSymmetryAtoms = _calculateSymmetryAtoms  # north_star spelling

    def _requireRatio(self):
        if not self.densityElectronRatio:
            raise RuntimeError("Failed to calculate densityElectronRatio, probably due to total aggregated electrons less than the minimum.")
        return self.densityElectronRatio

    # ------------------------------------------------------------------------------------------ blob statistics
    blobStatisticsHeader = ['distance_to_atom', 'sign', 'electrons_of_discrepancy', 'num_voxels', 'volume', 'chain',
                            'residue_number', 'residue_name', 'atom_name', 'atom_symmetry', 'atom_xyz', 'centroid_xyz']

    def calculateAtomSpecificBlobStatistics(self, blobList):
        """Nearest (symmetry) atom and electron count of every blob (pdb_eda/densityAnalysis.py:914-939): one
        brute-force float64 nearest-neighbour kernel over all blobs instead of one cdist call per blob."""
        symmetryAtoms = self.symmetryAtoms
        symmetryAtomCoords = self.symmetryAtomCoords
        densityElectronRatio = self._requireRatio()
        if not blobList:
            return []
        centroids = np.array([blob.centroid for blob in blobList], dtype=np.float64).reshape(-1, 3)
        idx, dist = _device.nearest_atom(centroids, np.asarray(symmetryAtomCoords, dtype=np.float64))
        idx, dist = idx.cpu().numpy(), dist.cpu().numpy()
        blobStats = []
        for blob, i, d in zip(blobList, idx.tolist(), dist):
            atom = symmetryAtoms[i]
            sign = '+' if blob.totalDensity >= 0 else '-'
            blobStats.append([d, sign, abs(blob.totalDensity / densityElectronRatio), len(blob), blob.volume, atom.parent.parent.id,
                              atom.parent.id[1], atom.parent.resname, atom.name, atom.symmetry, atom.coord, blob.centroid])
        return blobStats

    calcAtomBlobDists = calculateAtomSpecificBlobStatistics  # north_star spelling

    # ------------------------------------------------------------------------------------------ region density
    regionDensityHeader = ["actual_significant_regional_density", "num_electrons_actual_significant_regional_density"]
    atomRegionDensityHeader = ['model', 'chain', 'residue_number', 'residue_name', "atom_name", "occupancy"] + regionDensityHeader
    symmetryAtomRegionDensityHeader = ['model', 'chain', 'residue_number', 'residue_name', "atom_name", "symmetry", "atom_xyz",
                                       "fully_within_density_map"] + regionDensityHeader
    residueRegionDensityHeader = ['model', 'chain', 'residue_number', 'residue_name', "mean_occupancy"] + regionDensityHeader

    def _testRadius(self, atom, radius, useOptimizedRadii):
        resAtom = residueAtomName(atom)
        if useOptimizedRadii and resAtom in fullAtomNameMapAtomTypeGlobal:
            return radiiGlobal[fullAtomNameMapAtomTypeGlobal[resAtom]]
        return radius

    def _regionDensityRows(self, xyz, radii, groupStart, numSD):
        """Batched calculateRegionDensity: per group [sum of rho > cutoff over the union of spheres, / ratio], valid flag."""
        ratio = self._requireRatio()
        densityObj = self.densityObj
        densityCutoff = densityObj.meanDensity + numSD * densityObj.stdDensity
        out = utils.sphereSums(densityObj, xyz, radii, groupStart, densityCutoff, 0.0)
        return [[row[3], row[3] / ratio] for row in out], out[:, 6] != 0

    def calculateAtomRegionDensity(self, radius, numSD=1.5, type="", useOptimizedRadii=False):
        """Significant density within ``radius`` of every atom (pdb_eda/densityAnalysis.py:948-971), all atoms in one launch."""
        atoms = list(self.biopdbObj.get_atoms())
        if type:
            atoms = [atom for atom in atoms if atom.name == type]
        self._requireRatio()
        if not atoms:
            return []
        radii = np.array([self._testRadius(atom, radius, useOptimizedRadii) for atom in atoms], dtype=np.float64)
        rows, _ = self._regionDensityRows([atom.coord for atom in atoms], radii, None, numSD)
        return [[atom.parent.parent.parent.id, atom.parent.parent.id, atom.parent.id[1], atom.parent.resname, atom.name,
                 atom.get_occupancy()] + row for atom, row in zip(atoms, rows)]

    def calculateSymmetryAtomRegionDensity(self, radius, numSD=1.5, type="", useOptimizedRadii=False):
        """The same around every symmetry atom, with the fully-within-map flag (pdb_eda/densityAnalysis.py:973-999)."""
        atoms = self.symmetryAtoms
        if type:
            atoms = [atom for atom in atoms if atom.name == type]
        self._requireRatio()
        if not atoms:
            return []
        radii = np.array([self._testRadius(atom, radius, useOptimizedRadii) for atom in atoms], dtype=np.float64)
        rows, valid = self._regionDensityRows([np.asarray(atom.coord, dtype=np.float64) for atom in atoms], radii, None, numSD)
        return [[atom.parent.parent.parent.id, atom.parent.parent.id, atom.parent.id[1], atom.parent.resname, atom.name, atom.symmetry,
                 atom.coord, bool(ok)] + row for atom, row, ok in zip(atoms, rows, valid)]

    def calculateResidueRegionDensity(self, radius, numSD=1.5, type="", atomMask=None, useOptimizedRadii=False):
        """Significant density in the union of spheres around a residue's (masked) atoms
        (pdb_eda/densityAnalysis.py:1001-1035); all residues in one launch, one CTA per residue."""
        residues = list(self.biopdbObj.get_residues())
        if type:
            residues = [residue for residue in residues if residue.resname == type]
        self._requireRatio()
        kept, xyz, radii, start = [], [], [], [0]
        for residue in residues:
            atoms = [atom for atom in residue.get_atoms()
                     if not atomMask or residue.resname not in atomMask or atom.name in atomMask[residue.resname]]
            if atoms:
                kept.append((residue, np.mean([atom.get_occupancy() for atom in atoms])))
                xyz.extend(atom.coord for atom in atoms)
                radii.extend(self._testRadius(atom, radius, useOptimizedRadii) for atom in atoms)
                start.append(len(xyz))
        if not kept:
            return []
        rows, _ = self._regionDensityRows(xyz, np.asarray(radii, dtype=np.float64), np.asarray(start, dtype=np.int32), numSD)
        return [[residue.parent.parent.id, residue.parent.id, residue.id[1], residue.resname, meanOccupancy] + row
                for (residue, meanOccupancy), row in zip(kept, rows)]

    def calculateRegionDensity(self, xyzCoordList, radius, numSD=1.5, testValidCrs=False):
        """Region density of one list of points (pdb_eda/densityAnalysis.py:1037-1068)."""
        self._requireRatio()
        n = len(xyzCoordList)
        radii = np.asarray(radius[:n] if isinstance(radius, list) else [radius] * n, dtype=np.float64)
        if n == 0:
            raise IndexError("list index out of range")  # the reference indexes xyzCoords[0] (pdb_eda/ccp4.py:453)
        rows, valid = self._regionDensityRows(list(xyzCoordList)[:len(radii)], radii, np.array([0, len(radii)], dtype=np.int32), numSD)
        return (rows[0], bool(valid[0])) if testValidCrs else rows[0]

    # ------------------------------------------------------------------------------------------ region discrepancy
    regionDiscrepancyHeader = ["actual_abs_significant_regional_discrepancy", "num_electrons_actual_abs_significant_regional_discrepancy",
                               "expected_abs_significant_regional_discrepancy", "num_electrons_expected_abs_significant_regional_discrepancy",
                               "actual_significant_regional_discrepancy", "num_electrons_actual_significant_regional_discrepancy",
                               "actual_positive_significant_regional_discrepancy", "num_electrons_actual_positive_significant_regional_discrepancy",
                               "actual_negative_significant_regional_discrepancy", "num_electrons_actual_negative_significant_regional_discrepancy"]
    atomRegionDiscrepancyHeader = ['model', 'chain', 'residue_number', 'residue_name', "atom_name", "occupancy"] + regionDiscrepancyHeader
    symmetryAtomRegionDiscrepancyHeader = ['model', 'chain', 'residue_number', 'residue_name', "atom_name", "symmetry", "atom_xyz",
                                           "fully_within_density_map"] + regionDiscrepancyHeader
    residueRegionDiscrepancyHeader = ['model', 'chain', 'residue_number', 'residue_name', "mean_occupancy"] + regionDiscrepancyHeader

    def _regionDiscrepancyRows(self, xyz, radius, groupStart, numSD):
        """Batched calculateRegionDiscrepancy (pdb_eda/densityAnalysis.py:1160-1211): the ten values per group + valid flag."""
        ratio = self._requireRatio()
        diff = self.diffDensityObj
        cutoff = diff.meanDensity + numSD * diff.stdDensity
        n = len(xyz)
        out = utils.sphereSums(diff, xyz, np.full(n, radius, dtype=np.float64), groupStart, cutoff, -1.0 * cutoff)
        total_abs = diff.getTotalAbsDensity(cutoff)
        avg_abs_vox = total_abs / len(diff.densityArray)
        rows = []
        for row in out:
            pos, neg = row[3], row[5]
            actual = pos + neg
            actual_abs = abs(pos) + abs(neg)
            expected = avg_abs_vox * int(row[0])
            rows.append([actual_abs, actual_abs / ratio, expected, expected / ratio, actual, actual / ratio, pos, pos / ratio, neg, neg / ratio])
        return rows, out[:, 6] != 0

    def calculateAtomRegionDiscrepancies(self, radius, numSD=3.0, type=""):
        """pdb_eda/densityAnalysis.py:1081-1102, all atoms in one launch."""
        atoms = list(self.biopdbObj.get_atoms())
        if type:
            atoms = [atom for atom in atoms if atom.name == type]
        self._requireRatio()
        if not atoms:
            return []
        rows, _ = self._regionDiscrepancyRows([atom.coord for atom in atoms], radius, None, numSD)
        return [[atom.parent.parent.parent.id, atom.parent.parent.id, atom.parent.id[1], atom.parent.resname, atom.name,
                 atom.get_occupancy()] + row for atom, row in zip(atoms, rows)]

    def calculateSymmetryAtomRegionDiscrepancies(self, radius, numSD=3.0, type=""):
        """pdb_eda/densityAnalysis.py:1104-1128."""
        atoms = self.symmetryAtoms
        if type:
            atoms = [atom for atom in atoms if atom.name == type]
        self._requireRatio()
        if not atoms:
            return []
        rows, valid = self._regionDiscrepancyRows([np.asarray(atom.coord, dtype=np.float64) for atom in atoms], radius, None, numSD)
        return [[atom.parent.parent.parent.id, atom.parent.parent.id, atom.parent.id[1], atom.parent.resname, atom.name, atom.symmetry,
                 atom.coord, bool(ok)] + row for atom, row, ok in zip(atoms, rows, valid)]

    def calculateResidueRegionDiscrepancies(self, radius, numSD=3.0, type="", atomMask=None):
        """pdb_eda/densityAnalysis.py:1130-1158.  Mask semantics differ from the density variant: an atom is kept only
        if its residue name is in the mask and its name is listed."""
        residues = list(self.biopdbObj.get_residues())
        if type:
            residues = [residue for residue in residues if residue.resname == type]
        self._requireRatio()
        kept, xyz, start = [], [], [0]
        for residue in residues:
            atoms = [atom for atom in residue.get_atoms() if not atomMask or (residue.resname in atomMask and atom.name in atomMask[residue.resname])]
            if not atoms:
                raise IndexError("list index out of range")  # the reference fails on a residue left without atoms
            kept.append((residue, np.mean([atom.get_occupancy() for atom in atoms])))
            xyz.extend(atom.coord for atom in atoms)
            start.append(len(xyz))
        if not kept:
            return []
        rows, _ = self._regionDiscrepancyRows(xyz, radius, np.asarray(start, dtype=np.int32), numSD)
        return [[residue.parent.parent.id, residue.parent.id, residue.id[1], residue.resname, meanOccupancy] + row
                for (residue, meanOccupancy), row in zip(kept, rows)]

    def calculateRegionDiscrepancy(self, xyzCoordList, radius, numSD=3.0, testValidCrs=False):
        """Region discrepancy of one list of points (pdb_eda/densityAnalysis.py:1160-1211)."""
        self._requireRatio()
        if len(xyzCoordList) == 0:
            raise IndexError("list index out of range")
        rows, valid = self._regionDiscrepancyRows(list(xyzCoordList), radius, np.array([0, len(xyzCoordList)], dtype=np.int32), numSD)
        return (rows[0], bool(valid[0])) if testValidCrs else rows[0]

    # ------------------------------------------------------------------------------------------ F000
    def estimateF000(self):
        """Sum of electrons over the cell volume (pdb_eda/densityAnalysis.py:1214-1241); host arithmetic only."""
        if not elementElectronsGlobal:
            loadF000Parameters()
        totalElectrons = 0
        for atom in self.biopdbObj.get_atoms():
            fullAtomName = residueAtomName(atom)
            if fullAtomName in masterFullAtomNameMapElectronsGlobal:
                totalElectrons += masterFullAtomNameMapElectronsGlobal[fullAtomName]
            elif atom.element in elementElectronsGlobal:
                totalElectrons += elementElectronsGlobal[atom.element] + 1
        totalElectrons *= len(self.pdbObj.header.rotationMats)
        header = self.densityObj.header
        return totalElectrons / (header.unitVolume * header.nintervalX * header.nintervalY * header.nintervalZ)
