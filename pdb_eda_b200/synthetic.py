"""Synthetic CCP4 / PDB inputs with fixed seeds.

There is no network and no data set in the build and bench environments, so every test and benchmark input is
generated: little-endian mode-2 CCP4 files (``crsStart`` origin branch), a poly-ALA random-walk structure, a
2Fo-Fc map built from the atoms and an Fo-Fc map of smoothed noise.  The same bytes feed this package and the
reference (``oracle/_ref``) in the parity tests.
"""
import io
import struct

import numpy as np

from . import structure as _structure

# 12 symmetry operators of P6(5)22 in fractional coordinates (rotation rows + translation), International Tables.
_P6522_FRAC = [
    ([[1, 0, 0], [0, 1, 0], [0, 0, 1]], [0, 0, 0]),
    ([[0, -1, 0], [1, -1, 0], [0, 0, 1]], [0, 0, 2 / 3]),
    ([[-1, 1, 0], [-1, 0, 0], [0, 0, 1]], [0, 0, 1 / 3]),
    ([[-1, 0, 0], [0, -1, 0], [0, 0, 1]], [0, 0, 1 / 2]),
    ([[0, 1, 0], [-1, 1, 0], [0, 0, 1]], [0, 0, 1 / 6]),
    ([[1, -1, 0], [1, 0, 0], [0, 0, 1]], [0, 0, 5 / 6]),
    ([[0, 1, 0], [1, 0, 0], [0, 0, -1]], [0, 0, 2 / 3]),
    ([[1, -1, 0], [0, -1, 0], [0, 0, -1]], [0, 0, 0]),
    ([[-1, 0, 0], [-1, 1, 0], [0, 0, -1]], [0, 0, 1 / 3]),
    ([[0, -1, 0], [-1, 0, 0], [0, 0, -1]], [0, 0, 1 / 6]),
    ([[-1, 1, 0], [0, 1, 0], [0, 0, -1]], [0, 0, 1 / 2]),
    ([[1, 0, 0], [1, -1, 0], [0, 0, -1]], [0, 0, 5 / 6]),
]
_FRAC_OPS = {
    "P 1": [([[1, 0, 0], [0, 1, 0], [0, 0, 1]], [0, 0, 0])],
    "P 1 21 1": [([[1, 0, 0], [0, 1, 0], [0, 0, 1]], [0, 0, 0]), ([[-1, 0, 0], [0, 1, 0], [0, 0, -1]], [0, .5, 0])],
    "P 21 21 21": [([[1, 0, 0], [0, 1, 0], [0, 0, 1]], [0, 0, 0]), ([[-1, 0, 0], [0, -1, 0], [0, 0, 1]], [.5, 0, .5]),
                   ([[-1, 0, 0], [0, 1, 0], [0, 0, -1]], [0, .5, .5]), ([[1, 0, 0], [0, -1, 0], [0, 0, -1]], [.5, .5, 0])],
    "P 43 21 2": [([[1, 0, 0], [0, 1, 0], [0, 0, 1]], [0, 0, 0]), ([[-1, 0, 0], [0, -1, 0], [0, 0, 1]], [0, 0, .5]),
                  ([[0, -1, 0], [1, 0, 0], [0, 0, 1]], [.5, .5, .75]), ([[0, 1, 0], [-1, 0, 0], [0, 0, 1]], [.5, .5, .25]),
                  ([[-1, 0, 0], [0, 1, 0], [0, 0, -1]], [.5, .5, .75]), ([[1, 0, 0], [0, -1, 0], [0, 0, -1]], [.5, .5, .25]),
                  ([[0, 1, 0], [1, 0, 0], [0, 0, -1]], [0, 0, 0]), ([[0, -1, 0], [-1, 0, 0], [0, 0, -1]], [0, 0, .5])],
    "P 65 2 2": _P6522_FRAC,
}
SPACE_GROUP_NUMBER = {"P 1": 1, "P 1 21 1": 4, "P 21 21 21": 19, "P 43 21 2": 96, "P 65 2 2": 179}


def orthoMatrix(cell):
    """Orthogonalisation matrix of a cell (a, b, c, alpha, beta, gamma), float64 (Rupp p. 233)."""
    a, b, c, al, be, ga = cell
    ca, cb, cg = (np.cos(np.radians(x)) for x in (al, be, ga))
    sg = np.sin(np.radians(ga))
    v = np.sqrt(1 - ca * ca - cb * cb - cg * cg + 2 * ca * cb * cg)
    return np.array([[a, b * cg, c * cb], [0, b * sg, c * (ca - cb * cg) / sg], [0, 0, c * v / sg]], dtype=np.float64)


def cartesianOperators(spaceGroup, cell):
    """REMARK 290-style Cartesian 3x4 operators [O R O^-1 | O t], rounded like the PDB prints them (6 / 5 decimals)."""
    omat = orthoMatrix(cell)
    inv = np.linalg.inv(omat)
    ops = []
    for rot, trans in _FRAC_OPS[spaceGroup]:
        r = omat @ np.asarray(rot, dtype=np.float64) @ inv
        t = omat @ np.asarray(trans, dtype=np.float64)
        ops.append(np.concatenate((np.round(r, 6) + 0.0, (np.round(t, 5) + 0.0)[:, None]), axis=1))
    return ops


def ccp4Header(ncrs, cell, intervals, crsStart=(0, 0, 0), axisOrder=(1, 2, 3), spaceGroupNumber=1, stats=(0.0, 0.0, 0.0, 0.0)):
    """The 1,024 header bytes of a little-endian mode-2 CCP4 file (format string of pdb_eda/ccp4.py:149).
    ``ncrs`` = (columns, rows, sections); ``stats`` = (min, max, mean, rms)."""
    header = struct.pack("<10i6f3i3f3i27f4cifi", ncrs[0], ncrs[1], ncrs[2], 2, crsStart[0], crsStart[1], crsStart[2],
                         intervals[0], intervals[1], intervals[2], *[float(x) for x in cell], axisOrder[0], axisOrder[1],
                         axisOrder[2], float(stats[0]), float(stats[1]), float(stats[2]), spaceGroupNumber, 0, 0,
                         *([0.0] * 27), b"M", b"A", b"P", b" ", 0x00004144, float(stats[3]), 0)
    return header + b" " * 800


def ccp4Bytes(values, cell, intervals, crsStart=(0, 0, 0), axisOrder=(1, 2, 3), spaceGroupNumber=1):
    """A complete little-endian mode-2 CCP4 file.  ``values``: float32 array [section][row][column]."""
    values = np.ascontiguousarray(values, dtype="<f4")
    ns, nr, nc = values.shape
    stats = (values.min(), values.max(), values.mean(), values.std())
    return ccp4Header((nc, nr, ns), cell, intervals, crsStart, axisOrder, spaceGroupNumber, stats) + values.tobytes()


def ccp4Handle(values, cell, intervals, **kw):
    return io.BytesIO(ccp4Bytes(values, cell, intervals, **kw))


# ALA backbone + CB offsets from CA (Angstrom), an idealised residue frame
_ALA_ATOMS = (("N", (-0.53, 1.36, 0.0), "N"), ("CA", (0.0, 0.0, 0.0), "C"), ("C", (1.53, 0.0, 0.0), "C"),
              ("O", (2.15, -1.06, 0.0), "O"), ("CB", (-0.53, -0.77, -1.21), "C"))
ALA_ELECTRONS = {"ALA_N": 8.0, "ALA_CA": 7.0, "ALA_C": 6.0, "ALA_O": 8.0, "ALA_CB": 9.0}


def defaultParams():
    """A small parameter dictionary in the layout of the reference's conf/optimized_params.json
    (pdb_eda/densityAnalysis.py:32-46) covering the synthetic poly-ALA atoms."""
    types = {"ALA_N": "N.N.8#C.N.7.SING_H.N.1.SING", "ALA_CA": "C.N.7#C.N.6.SING_C.N.9.SING_H.N.1.SING_N.N.8.SING",
             "ALA_C": "C.N.6#C.N.7.SING_O.N.8.DOUB", "ALA_O": "O.N.8#C.N.6.DOUB",
             "ALA_CB": "C.N.9#C.N.7.SING_H.N.1.SING_H.N.1.SING_H.N.1.SING"}
    radii = {types["ALA_N"]: 0.78, types["ALA_CA"]: 0.72, types["ALA_C"]: 0.66, types["ALA_O"]: 0.81, types["ALA_CB"]: 0.84}
    return {"radii": radii, "slopes": {t: 0.0 for t in radii},
            "bonded_atoms": {"ALA_N": ["ALA_CA"], "ALA_CA": ["ALA_N", "ALA_C", "ALA_CB"], "ALA_C": ["ALA_CA", "ALA_O"],
                             "ALA_O": ["ALA_C"], "ALA_CB": ["ALA_CA"]},
            "full_atom_name_map_electrons": dict(ALA_ELECTRONS), "full_atom_name_map_atom_type": types,
            "leaving_atoms": []}


def _random_rotation(rng):
    q = rng.normal(size=4)
    q /= np.linalg.norm(q)
    w, x, y, z = q
    return np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w)],
                     [2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w)],
                     [2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)]])


def polyAlaStructure(nResidues, boxLo, boxHi, seed=1, structureId="synth", residuesPerChain=250, hetero=0):
    """A poly-ALA chain whose CA atoms random-walk (3.8 A steps) inside [boxLo, boxHi]; coordinates rounded to
    3 decimals and stored as float32, occupancy 1.0, B in U(10, 40).  ``hetero`` waters are appended."""
    rng = np.random.default_rng(seed)
    lo = np.asarray(boxLo, dtype=np.float64) + 3.0
    hi = np.asarray(boxHi, dtype=np.float64) - 3.0
    st = _structure.Structure(structureId)
    st.header["resolution"] = 2.0
    model = st.add(_structure.Model(0))
    pos = lo + (hi - lo) * rng.uniform(0.3, 0.7, 3)
    chain = None
    for k in range(nResidues):
        if k % residuesPerChain == 0:
            chain = model.add(_structure.Chain(chr(ord("A") + (k // residuesPerChain) % 26)))
        for _ in range(64):
            step = rng.normal(size=3)
            step *= 3.8 / np.linalg.norm(step)
            if np.all(pos + step > lo) and np.all(pos + step < hi):
                pos = pos + step
                break
        rot = _random_rotation(rng)
        res = chain.add(_structure.Residue((" ", k + 1, " "), "ALA"))
        for name, off, element in _ALA_ATOMS:
            coord = np.round(pos + rot @ np.asarray(off), 3).astype(np.float32)
            res.add(_structure.Atom(name, coord, round(float(rng.uniform(10, 40)), 2), 1.0, element=element))
    if hetero:
        chain = model.add(_structure.Chain("W"))
        for k in range(hetero):
            res = chain.add(_structure.Residue(("W", k + 1, " "), "HOH"))
            coord = np.round(lo + (hi - lo) * rng.uniform(0, 1, 3), 3).astype(np.float32)
            res.add(_structure.Atom("O", coord, 30.0, 1.0, element="O"))
    return st


def _gaussian_blur_periodic(vol, sigma):
    from scipy import ndimage
    return ndimage.gaussian_filter(vol, sigma=sigma, mode="wrap")


def mapPair(structure, n, cell, seed=7, electrons=None, sigma=1.2, crsStart=(0, 0, 0), axisOrder=(1, 2, 3)):
    """(2Fo-Fc, Fo-Fc) float32 volumes [section][row][column] on an n-interval grid covering one cell.

    2Fo-Fc = electrons deposited at each atom's nearest grid point, Gaussian-smoothed (sigma voxels), plus 2 % noise;
    Fo-Fc = Gaussian-smoothed N(0,1) noise scaled to 0.1 std(2Fo-Fc).  Orthogonal or skewed cells.
    ``n`` = (nx, ny, nz) intervals; the stored grid holds the whole cell starting at crsStart.
    """
    rng = np.random.default_rng(seed)
    nx, ny, nz = n
    electrons = electrons or ALA_ELECTRONS
    inv = np.linalg.inv(orthoMatrix(cell))
    vol = np.zeros((nx, ny, nz), dtype=np.float64)  # indexed [x][y][z] on the unit cell grid
    for atom in structure.get_atoms():
        frac = inv @ atom.coord.astype(np.float64)
        g = np.rint(frac * (nx, ny, nz)).astype(int) % (nx, ny, nz)
        vol[g[0], g[1], g[2]] += electrons.get(atom.parent.resname + "_" + atom.name, 6.0)
    vol = _gaussian_blur_periodic(vol, sigma)
    vol += 0.02 * vol.std() * rng.standard_normal(vol.shape)
    diff = _gaussian_blur_periodic(rng.standard_normal(vol.shape), sigma)
    diff *= 0.1 * vol.std() / diff.std()
    return _to_crs(vol, n, crsStart, axisOrder), _to_crs(diff, n, crsStart, axisOrder)


def _to_crs(volXYZ, n, crsStart, axisOrder, ncrs=None):
    """Re-index a unit-cell volume [x][y][z] into the stored [section][row][column] array of a CCP4 file whose
    column/row/section axes carry axisOrder (1=x, 2=y, 3=z) and start at crsStart; periodic."""
    axes = [a - 1 for a in axisOrder]  # xyz axis carried by column, row, section
    if ncrs is None:
        ncrs = [n[axes[0]], n[axes[1]], n[axes[2]]]
    idx = [(np.arange(ncrs[k]) + crsStart[k]) % n[axes[k]] for k in range(3)]
    # out[s][r][c] = vol[x][y][z] with the xyz index of each crs axis
    sel = [None, None, None]
    for k in range(3):
        sel[axes[k]] = idx[k]
    shape_of = {axes[0]: 2, axes[1]: 1, axes[2]: 0}  # xyz axis -> position in (s, r, c)
    grids = [None, None, None]
    for xyz_axis in range(3):
        shp = [1, 1, 1]
        shp[shape_of[xyz_axis]] = len(sel[xyz_axis])
        grids[xyz_axis] = sel[xyz_axis].reshape(shp)
    return np.ascontiguousarray(volXYZ[grids[0], grids[1], grids[2]], dtype=np.float32)


def smoothNoiseMap(n, seed=4, sigma=1.5):
    """Fo-Fc-like volume of smoothed unit-variance noise, float32 [n][n][n] (config 4 style, host version)."""
    rng = np.random.default_rng(seed)
    vol = _gaussian_blur_periodic(rng.standard_normal((n, n, n)).astype(np.float32), sigma)
    vol /= vol.std()
    return vol.astype(np.float32)


def smoothNoiseMapDevice(n, seed=4, passes=2, device="cuda"):
    """Large smoothed-noise volumes generated directly in HBM (bench input only): iterated periodic 3-point box
    filters along each axis of N(0,1) noise, normalised to unit variance.  Returns a float32 CUDA tensor [n][n][n]."""
    import torch
    gen = torch.Generator(device=device)
    gen.manual_seed(seed)
    vol = torch.randn((n, n, n), generator=gen, device=device, dtype=torch.float32)
    for _ in range(passes):
        for axis in range(3):
            vol = (vol + torch.roll(vol, 1, axis) + torch.roll(vol, -1, axis)) / 3.0
    vol = vol / vol.std()
    return vol.contiguous()


# ---------------------------------------------------------------------------------------------------- structure pools
# BASELINE.json config 3: many synthetic structures of mixed size and space group.  Everything below is input synthesis
# (numpy + torch on the device), none of it is on the measured path.
POOL_SIZES = (64, 96, 128, 192, 256)
POOL_SPACE_GROUPS = ("P 1", "P 1 21 1", "P 21 21 21", "P 43 21 2", "P 65 2 2")


def poolSpec(nStructures, seed=3, sizes=POOL_SIZES, spaceGroups=POOL_SPACE_GROUPS, atomsPerVoxel=1.0 / 350.0):
    """The pool as a list of dicts (index, n, cell, spaceGroup, residues, seed): grid sizes and space groups drawn with
    ``seed``, 0.5 A grid, atoms ~ n^3 / 350.  P 65 2 2 gets a hexagonal cell (gamma = 120)."""
    rng = np.random.default_rng(seed)
    spec = []
    for k in range(nStructures):
        n = int(sizes[int(rng.integers(len(sizes)))])
        sg = spaceGroups[int(rng.integers(len(spaceGroups)))]
        edge = 0.5 * n
        cell = (edge, edge, edge, 90.0, 90.0, 120.0 if sg == "P 65 2 2" else 90.0)
        volumeFraction = np.sin(np.radians(cell[5]))
        residues = max(int(round(n ** 3 * atomsPerVoxel * volumeFraction / 5.0)), 4)
        spec.append({"index": k, "n": n, "cell": cell, "spaceGroup": sg, "residues": residues, "seed": 1000 + seed * 100003 + k})
    return spec


def fastPolyAla(nResidues, cell, seed):
    """Vectorised poly-ALA generator: CA random walk (3.8 A steps, reflected into a box inscribed in the cell), a random
    rotation of the residue frame per residue, coordinates rounded to 3 decimals as float32.
    Returns (coords32 (5 n, 3), bfactor (5 n))."""
    rng = np.random.default_rng(seed)
    omat = orthoMatrix(cell)
    if cell[5] != 90:
        corners = np.array([omat @ np.array([0.28, 0.3, 0.08]), omat @ np.array([0.55, 0.7, 0.92])])
        lo, hi = corners.min(axis=0) + 2.5, corners.max(axis=0) - 2.5
        lo[0], hi[0] = omat[0, 0] * 0.3, omat[0, 0] * 0.55      # stay inside the sheared x range for every y of the box
    else:
        lo, hi = np.full(3, 3.0), np.asarray(cell[:3], dtype=np.float64) - 3.0
    steps = rng.normal(size=(nResidues, 3))
    steps *= 3.8 / np.linalg.norm(steps, axis=1, keepdims=True)
    span = hi - lo
    walk = (span * rng.uniform(0.3, 0.7, 3)) + np.cumsum(steps, axis=0)
    folded = np.mod(walk, 2 * span)
    ca = lo + np.where(folded > span, 2 * span - folded, folded)           # triangle wave: reflection at the box faces
    q = rng.normal(size=(nResidues, 4))
    q /= np.linalg.norm(q, axis=1, keepdims=True)
    w, x, y, z = q.T
    rot = np.stack([np.stack([1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w)], axis=1),
                    np.stack([2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w)], axis=1),
                    np.stack([2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)], axis=1)], axis=1)
    offs = np.array([o for _, o, _ in _ALA_ATOMS], dtype=np.float64)           # (5, 3)
    coords = ca[:, None, :] + np.einsum("nij,aj->nai", rot, offs)
    coords32 = np.round(coords.reshape(-1, 3), 3).astype(np.float32)
    bfactor = np.round(rng.uniform(10, 40, 5 * nResidues), 2)
    return coords32, bfactor


def polyAlaTable(coords32, bfactor, params):
    """``cloudBatch.AtomTable`` of a poly-ALA chain given as arrays (the same rows AtomTable.fromStructure would yield)."""
    from .cloudBatch import AtomTable, _distinctRows
    names = ["ALA_" + a for a, _, _ in _ALA_ATOMS]
    n = len(coords32)
    nres = n // 5
    local = np.tile(np.arange(5, dtype=np.int32), nres)
    bondedOf = np.zeros(5, dtype=np.uint64)
    for k, name in enumerate(names):
        for other in params["bonded_atoms"].get(name, ()):
            if other in names:
                bondedOf[k] |= np.uint64(1 << names.index(other))
    supported = _distinctRows(coords32) == n
    return AtomTable(coords32, names, local.copy(), np.ones(n), bfactor, np.repeat(np.arange(nres, dtype=np.int32), 5), local,
                     np.tile(bondedOf, nres), nres, None, supported, "" if supported else "atoms with identical coordinates")


def densityMapDevice(coords32, electronsPerAtom, n, cell, seed, sigma=1.2, device="cuda"):
    """2Fo-Fc-like volume generated in HBM: electrons deposited at each atom's nearest grid point, periodic Gaussian smoothing
    (sigma voxels), 2 % noise -- the recipe of ``mapPair`` for column/row/section = x/y/z.  float32 CUDA tensor [n][n][n]."""
    import torch
    inv = np.linalg.inv(orthoMatrix(cell))
    frac = coords32.astype(np.float64) @ inv.T
    g = np.rint(frac * n).astype(np.int64) % n
    flat = torch.from_numpy((g[:, 2] * n + g[:, 1]) * n + g[:, 0]).to(device)
    vol = torch.zeros(n * n * n, dtype=torch.float32, device=device)
    vol.index_add_(0, flat, torch.from_numpy(np.asarray(electronsPerAtom, dtype=np.float32)).to(device))
    vol = vol.view(n, n, n)
    radius = int(np.ceil(4 * sigma))
    ks = np.exp(-0.5 * (np.arange(-radius, radius + 1) / sigma) ** 2)
    ks /= ks.sum()
    for axis in range(3):
        acc = torch.zeros_like(vol)
        for k, wgt in zip(range(-radius, radius + 1), ks):
            acc.add_(torch.roll(vol, k, axis), alpha=float(wgt))
        vol = acc
    gen = torch.Generator(device=device)
    gen.manual_seed(int(seed))
    vol = vol + 0.02 * vol.std() * torch.randn(vol.shape, generator=gen, device=device, dtype=torch.float32)
    return vol.contiguous()


def buildPoolEntries(spec, params, device="cuda"):
    """(pool index, cloudBatch.MapRef, cloudBatch.AtomTable) for every structure of ``spec`` (see ``poolSpec``): structure
    and 2Fo-Fc map synthesised on the spot, map resident on ``device``, densityCutoff = mean + 1.5 sigma from the library's
    own reduction (pdb_eda/densityAnalysis.py:131)."""
    from . import _device, ccp4
    from .cloudBatch import MapRef
    electronsOf = np.array([ALA_ELECTRONS["ALA_" + a] for a, _, _ in _ALA_ATOMS])
    entries = []
    for sp in spec:
        n, cell = sp["n"], sp["cell"]
        coords, bf = fastPolyAla(sp["residues"], cell, sp["seed"])
        table = polyAlaTable(coords, bf, params)
        rho = densityMapDevice(coords, np.tile(electronsOf, sp["residues"]), n, cell, sp["seed"] + 1, device=device)
        hdr = ccp4.DensityHeader.fromFileHeader(ccp4Header((n, n, n), cell, (n, n, n)))
        dmap = _device.DeviceMap(_device.geom_from_header(hdr), rho.reshape(-1))
        mean, std = dmap.mean_std()
        entries.append((sp["index"], MapRef(dmap, mean + 1.5 * std, hdr.unitVolume, "s%05d" % sp["index"]), table))
    return entries


def poolCosts(spec):
    """Cost estimate per structure for longest-first sharding: cloud voxels scale with the atom count."""
    return [5.0 * sp["residues"] for sp in spec]
