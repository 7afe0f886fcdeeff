"""GPU tier: the CUDA library (through the C ABI) against the CPU oracle on seeded random inputs, beyond the sizes
and shapes the golden vectors cover: ragged groups, empty inputs, very large and very small radii (every kernel
path: tabulated / generic, bitmap claims / earlier-atom claims, boxes wider than a warp), the reference's own
known-answer scenarios (planted cubes, wrap semantics, merge + testOverlap: tests/test_ccp4.py:68-131 of the
reference), and size-independent properties on large maps."""
import io

import numpy as np
import pytest

import cases
import golden_checks as gc

pytestmark = pytest.mark.gpu


def _impls(name, seed=11):
    from impl_cuda import CudaImpl
    from impl_oracle import OracleImpl
    from pdb_eda_b200 import ccp4
    data, _ = cases.make_case(name, seed)
    dm = ccp4.parse(io.BytesIO(data), name)
    return dm, CudaImpl(dm), OracleImpl(dm)


def _cmp_sums(a, b):
    for col in (0, 2, 4, 6):
        assert np.array_equal(a[:, col], b[:, col]), col
    for col in (1, 3, 5):
        gc.close(a[:, col], b[:, col], atol=1e-10)


@pytest.mark.parametrize("name", gc.CASES)
def test_union_ragged_groups(name):
    dm, cuda, orc = _impls(name)
    rng = np.random.default_rng(5)
    sizes = rng.integers(0, 9, 25)
    sizes[3] = 0
    start = np.concatenate(([0], np.cumsum(sizes))).astype(np.int32)
    n = int(start[-1])
    # atoms of a group lie close together (like a residue)
    centres = cases.random_atoms(dm, len(sizes), seed=6).astype(np.float64)
    xyz = np.concatenate([centres[k] + rng.uniform(-2.5, 2.5, (s, 3)) for k, s in enumerate(sizes)]) if n else np.zeros((0, 3))
    xyz = np.round(xyz, 3).astype(np.float32).astype(np.float64)
    radii = rng.uniform(1.0, 3.5, n).astype(np.float32)
    m, s = cuda.mean_std()
    cut = m + 1.0 * s
    _cmp_sums(cuda.sphere_sums(xyz, radii, start, cut, -cut), orc.sphere_sums(xyz, radii, start, cut, -cut))


@pytest.mark.parametrize("name,radius", [("ortho", 9.0), ("ortho", 12.5), ("perm", 12.5), ("hex", 6.0)])
def test_union_wide_boxes(name, radius):
    """Boxes wider than a warp (two column passes), and groups whose bounding box exceeds the shared bitmap."""
    dm, cuda, orc = _impls(name)
    rng = np.random.default_rng(8)
    centres = cases.random_atoms(dm, 2, seed=9).astype(np.float64)
    xyz = np.concatenate([c + rng.uniform(-2, 2, (3, 3)) for c in centres])
    xyz = np.round(xyz, 3).astype(np.float32).astype(np.float64)
    start = np.array([0, 3, 6], dtype=np.int32)
    radii = np.full(6, radius, dtype=np.float32)
    _cmp_sums(cuda.sphere_sums(xyz, radii, start, 0.5, -0.5), orc.sphere_sums(xyz, radii, start, 0.5, -0.5))
    _cmp_sums(cuda.sphere_sums(xyz[:2], radii[:2], None, 0.5, -0.5), orc.sphere_sums(xyz[:2], radii[:2], None, 0.5, -0.5))


@pytest.mark.parametrize("name", ["ortho", "perm"])
def test_union_borderline_chords(name):
    """Atoms on grid points (and within a few ulp / 1e-4 columns of them) with radii that are whole multiples of the
    grid spacing and Pythagorean combinations of it: many voxels lie at distance == r or within rounding of it, and
    the chord ends of the union kernel's float32 interval guess fall on grid points -- the cases where the guess must
    not be trusted and the exact float64 predicate decides."""
    dm, cuda, orc = _impls(name)
    h = dm.header
    rng = np.random.default_rng(21)
    gl = np.array([h.gridLength[i] for i in range(3)], dtype=np.float64)
    origin = np.array([h.origin[i] for i in range(3)], dtype=np.float64)
    base = []
    for _ in range(40):
        idx = rng.integers(8, 30, 3)
        p = origin + idx * gl
        jitter = rng.choice([0.0, 1e-7, -1e-7, 5e-5, -5e-5, 4e-4, -4e-4, 6e-4], 3) * gl
        base.append(p + jitter)
    xyz = np.array(base, dtype=np.float64)  # float64 coordinates (as symmetry images are): keeps the tiny offsets
    g0 = float(gl.min())
    rads = [g0 * k for k in (1, 2, 3, 5, 7)] + [g0 * np.sqrt(k) for k in (2, 3, 5, 13, 25, 50)]
    radii = np.array([rads[i % len(rads)] for i in range(len(xyz))], dtype=np.float32)
    for per in (1, 4):
        start = np.arange(0, len(xyz) + 1, per, dtype=np.int32)
        m, s = cuda.mean_std()
        cut = m + 0.5 * s
        _cmp_sums(cuda.sphere_sums(xyz, radii, start, cut, -cut), orc.sphere_sums(xyz, radii, start, cut, -cut))


@pytest.mark.parametrize("name", ["ortho", "tric"])
def test_degenerate_radii_and_empty_batches(name):
    dm, cuda, orc = _impls(name)
    xyz = cases.random_atoms(dm, 6, seed=3).astype(np.float64)
    radii = np.array([0.0, 1e-3, 0.2, 0.25, 0.5, 30.0 if name == "ortho" else 3.0], dtype=np.float32)
    _cmp_sums(cuda.sphere_sums(xyz, radii, None, 0.0, 0.0), orc.sphere_sums(xyz, radii, None, 0.0, 0.0))
    crs_c, off_c = cuda.sphere_lists(xyz[:5], radii[:5], 0.0)
    crs_o, off_o = orc.sphere_lists(xyz[:5], radii[:5], 0.0)
    assert np.array_equal(off_c, off_o) and np.array_equal(crs_c, crs_o)
    # empty batches
    assert cuda.sphere_sums(np.zeros((0, 3)), np.zeros(0, np.float32), None, 0.0, 0.0).shape == (0, 8)
    crs_e, off_e = cuda.sphere_lists(np.zeros((0, 3)), np.zeros(0, np.float32), 0.0)
    assert len(crs_e) == 0 and list(off_e) == [0]


@pytest.mark.parametrize("name", ["ortho", "perm", "hex"])
def test_sphere_lists_and_clouds_large_radius(name):
    """Per-atom clusters (findAberrantBlobs) at a region-size radius: many clusters per atom."""
    dm, cuda, orc = _impls(name)
    xyz = cases.random_atoms(dm, 10, seed=21).astype(np.float64)
    radii = np.full(10, 3.5, dtype=np.float32)
    m, s = cuda.mean_std()
    for cut in (m + 1.2 * s, -(m + 1.2 * s)):
        ca, cb = cuda.sphere_lists(xyz, radii, cut), orc.sphere_lists(xyz, radii, cut)
        assert np.array_equal(ca[1], cb[1]) and np.array_equal(ca[0], cb[0])
        a, b = cuda.sphere_clouds(xyz, radii, cut), orc.sphere_clouds(xyz, radii, cut)
        assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1])
        gc.close(a[2], b[2])


def _planted(shape=(60, 50, 40), intervals=(112, 112, 192), cell=(56.0, 56.0, 96.0, 90, 90, 90), crsStart=(-7, 11, 5),
             axisOrder=(2, 1, 3)):
    from pdb_eda_b200 import ccp4, synthetic
    values = np.zeros(shape, dtype=np.float32)
    data = synthetic.ccp4Bytes(values, cell, intervals, crsStart=crsStart, axisOrder=axisOrder)
    return ccp4.parse(io.BytesIO(data), "kat")


def test_kat_planted_cubes_and_merge():
    """The reference's own scenarios (tests/test_ccp4.py:75-131): 2^3 and 3^3 cubes of +-1; merge of an 8-cube
    with an 8-cube plus a bridging voxel -> 17 voxels whose centroid is the centre."""
    dm = _planted()
    h = dm.header
    d = dm.density
    d[10:12, 10:12, 10:12] = 1.0       # 2^3 green
    d[20:23, 20:23, 20:23] = 1.0       # 3^3 green
    d[30:32, 5:7, 5:7] = -1.0          # 2^3 red
    d[5:8, 30:33, 30:33] = -1.0        # 3^3 red
    green = dm.createFullBlobList(0.5)
    red = dm.createFullBlobList(-0.5)
    assert [len(b.crsList) for b in green] == [8, 27] and [len(b.crsList) for b in red] == [8, 27]
    gc.close([b.totalDensity for b in green], [8.0, 27.0])
    gc.close([b.totalDensity for b in red], [-8.0, -27.0])
    gc.close([b.volume for b in green], [8 * h.unitVolume, 27 * h.unitVolume])
    centre = np.mean([h.crs2xyzCoord([c, r, s]) for c in (10, 11) for r in (10, 11) for s in (10, 11)], axis=0)
    gc.close(green[0].centroid, centre, rtol=1e-9, atol=1e-9)
    gc.close(green[0].coordCenter, centre, rtol=1e-9, atol=1e-9)
    # green and red from one pass agree with the two single-sign calls
    g2, r2 = dm.createFullBlobLists(0.5, -0.5)
    assert all(a == b for a, b in zip(green, g2)) and all(a == b for a, b in zip(red, r2))
    assert dm.createFullBlobList(0) is None
    # merge + overlap
    dm2 = _planted()
    dm2.density[10:12, 10:12, 10:12] = 1.0
    dm2.density[13:15, 10:12, 10:12] = 1.0
    blobs = dm2.createFullBlobList(0.5)
    assert len(blobs) == 2 and not blobs[0].testOverlap(blobs[1])
    dm2.density[12, 11, 11] = 1.0       # the bridging voxel joins the cubes (through the tracked host array)
    joined = dm2.createFullBlobList(0.5)
    assert len(joined) == 1 and len(joined[0].crsList) == 17
    a, b = blobs
    b.crsList = set(b.crsList) | {(11, 11, 12)}
    assert a.testOverlap(b)
    a.merge(b)
    assert len(a.crsList) == 17 and a == joined[0]


def test_kat_wrap_semantics():
    """getPointDensityFromCrs repeats after the interval, is 0 where the cell is not covered (tests/test_ccp4.py:68-72)."""
    from pdb_eda_b200 import cutils
    dm = _planted(shape=(20, 20, 20), intervals=(32, 32, 32), cell=(16.0, 16.0, 16.0, 90, 90, 90), crsStart=(0, 0, 0),
                  axisOrder=(1, 2, 3))
    dm.density[3, 4, 19] = 7.0
    dm.density[0, 0, 0] = 5.0
    assert cutils.getPointDensityFromCrs(dm, [19, 4, 3]) == 7.0
    assert cutils.getPointDensityFromCrs(dm, [-13, 4, 3]) == 7.0       # -13 + 32 = 19
    assert cutils.getPointDensityFromCrs(dm, [32, 32, 32]) == 5.0      # one interval further
    assert cutils.getPointDensityFromCrs(dm, [20, 0, 0]) == 0 and not cutils.testValidCrs(dm, [20, 0, 0])
    assert cutils.getPointDensityFromCrs(dm, [-1, 0, 0]) == 0 and not cutils.testValidCrs(dm, [31, 0, 0])
    assert cutils.testValidCrsList(dm, [(0, 0, 0), (32, 19, -32)]) and not cutils.testValidCrsList(dm, [(0, 0, 0), (25, 0, 0)])


@pytest.mark.parametrize("n", [128, 200])
def test_blobs_on_larger_maps(n):
    """Full blob labelling against the oracle at sizes the reference itself cannot cluster (O(N^2) cdist)."""
    from pdb_eda_b200 import ccp4, synthetic
    from impl_cuda import CudaImpl
    from impl_oracle import OracleImpl
    vol = synthetic.smoothNoiseMap(n, seed=n, sigma=1.5)
    nc = n - 8  # stored columns differ from the interval: exercises the non-vectorised path on odd sizes too
    data = synthetic.ccp4Bytes(vol[:, :, :nc], (n * 0.4,) * 3 + (90, 90, 90), (n, n, n))
    dm = ccp4.parse(io.BytesIO(data), "big")
    cuda, orc = CudaImpl(dm), OracleImpl(dm)
    for (c1, l1, s1), (c2, l2, s2) in zip(cuda.full_blobs(2.8, -2.8), orc.full_blobs(2.8, -2.8)):
        assert len(c1) > 1000
        assert np.array_equal(c1, c2) and np.array_equal(l1, l2)
        gc.close(s1, s2, rtol=1e-9, atol=1e-9)


def test_blob_properties_at_scale():
    """Size-independent properties on a 384^3 map (BASELINE.json config 2 size): labels are canonical (a blob's
    number is the rank of its first voxel), every blob is a single 26-connected component, blobs are mutually
    non-adjacent, and the per-blob counts add up."""
    import torch
    from pdb_eda_b200 import _device, ccp4, synthetic
    n = 384
    vol = synthetic.smoothNoiseMapDevice(n, seed=4)
    hdr = ccp4.DensityHeader.fromFileHeader(synthetic.ccp4Header((n, n, n), (192.0,) * 3 + (90, 90, 90), (n, n, n)))
    dev = _device.DeviceMap(_device.geom_from_header(hdr), vol.reshape(-1))
    m, s = dev.mean_std()
    green, red = dev.blob_label(m + 3 * s, -(m + 3 * s))
    for part in (green, red):
        crs, label, stats = part["crs"], part["label"].long(), part["stats"]
        nb = part["n_blobs"]
        assert part["n_voxels"] > 10000 and nb > 100
        key = (crs[:, 0].long() * n + crs[:, 1].long()) * n + crs[:, 2].long()
        assert bool((key[1:] > key[:-1]).all())                       # createFullCrsList order
        first = torch.full((nb,), len(label), dtype=torch.long, device=label.device)
        first.scatter_reduce_(0, label, torch.arange(len(label), device=label.device), reduce="amin")
        assert bool((first[1:] > first[:-1]).all())                   # blob b's first voxel precedes blob b+1's
        assert torch.equal(torch.bincount(label, minlength=nb).double(), stats[:, 0])
        # dense check of connectivity: no two voxels of different blobs are 26-adjacent
        dense = torch.full((n + 2, n + 2, n + 2), -1, dtype=torch.int32, device=label.device)
        dense[crs[:, 0].long() + 1, crs[:, 1].long() + 1, crs[:, 2].long() + 1] = label.int()
        core = dense[1:-1, 1:-1, 1:-1]
        for dc in (-1, 0, 1):
            for dr in (-1, 0, 1):
                for ds in (-1, 0, 1):
                    nb_lab = dense[1 + dc:n + 1 + dc, 1 + dr:n + 1 + dr, 1 + ds:n + 1 + ds]
                    both = (core >= 0) & (nb_lab >= 0)
                    assert bool((core[both] == nb_lab[both]).all())
        # every blob is connected: re-clustering its voxels as an arbitrary list gives the same partition
        lab2, ncl = _device.cluster_crs(crs)
        assert ncl == nb and torch.equal(lab2.long(), label)


def test_blob_dense_foreground_beyond_the_shared_rank_table():
    """More than 2 M foreground voxels (4,096 ranking chunks): the sparse stage ranks the roots through the global
    chunk prefix (one more grid barrier) instead of the per-block shared-memory table.  Checked against
    scipy.ndimage.label (26-connectivity) renumbered by first voxel in createFullCrsList order."""
    import scipy.ndimage as ndi
    from pdb_eda_b200 import _device, ccp4, synthetic
    n = 176
    import torch
    vol = np.ascontiguousarray(synthetic.smoothNoiseMap(n, seed=31, sigma=1.0), dtype=np.float32)
    hdr = ccp4.DensityHeader.fromFileHeader(synthetic.ccp4Header((n, n, n), (88.0,) * 3 + (90, 90, 90), (n, n, n)))
    dev = _device.DeviceMap(_device.geom_from_header(hdr), torch.from_numpy(vol.reshape(-1)).cuda())
    m, s = dev.mean_std()
    cut = float(np.float32(m + 0.3 * s))
    parts = dev.blob_label(cut, -cut, cap_voxels=3_000_000, cap_blobs=1_000_000)
    assert sum(p["n_voxels"] for p in parts) > 4096 * 512
    crs_order = np.ascontiguousarray(vol.transpose(2, 1, 0))  # [c][r][s]: C order = createFullCrsList order
    for part, mask in zip(parts, (crs_order >= np.float32(cut), crs_order <= np.float32(-cut))):
        lab, nlab = ndi.label(mask, structure=np.ones((3, 3, 3), dtype=bool))
        flat = lab[mask]
        assert part["n_voxels"] == flat.size and part["n_blobs"] == nlab
        _, first = np.unique(flat, return_index=True)
        remap = np.empty(nlab + 1, dtype=np.int64)
        remap[1 + np.argsort(first)] = np.arange(nlab)
        assert np.array_equal(part["label"].cpu().numpy().astype(np.int64), remap[flat])
        key = np.flatnonzero(mask.reshape(-1))
        crs = part["crs"].cpu().numpy().astype(np.int64)
        assert np.array_equal((crs[:, 0] * n + crs[:, 1]) * n + crs[:, 2], key)
        assert np.array_equal(np.bincount(remap[flat], minlength=nlab), part["stats"][:, 0].cpu().numpy().astype(np.int64))


def test_sphere_properties_at_scale():
    """BASELINE.json config 2 size (384^3, 40,000 atoms / 8,000 residues): size-independent properties of the sphere
    kernels plus an oracle check on a random sample of atoms and residues of the full-size problem."""
    import torch
    from oracle import orc
    from pdb_eda_b200 import _blas, _device, ccp4, synthetic
    n, nres = 384, 8000
    cell = n * 0.5
    vol = synthetic.smoothNoiseMapDevice(n, seed=11)
    hdr = ccp4.DensityHeader.fromFileHeader(synthetic.ccp4Header((n, n, n), (cell,) * 3 + (90, 90, 90), (n, n, n)))
    dev = _device.DeviceMap(_device.geom_from_header(hdr), vol.reshape(-1))
    rng = np.random.default_rng(12)
    ca = rng.uniform(-2, cell + 2, (nres, 3))                                   # some residues hang over the cell edge (wrap)
    offs = np.array([o for _, o, _ in synthetic._ALA_ATOMS])
    xyz = np.round((ca[:, None, :] + offs[None, :, :]).reshape(-1, 3), 3).astype(np.float32).astype(np.float64)
    start = np.arange(0, 5 * nres + 1, 5, dtype=np.int32)
    r35 = np.full(len(xyz), 3.5, dtype=np.float32)
    per_atom = dev.sphere_sums(xyz, r35, None, 1.0, -1.0).cpu().numpy()
    per_res = dev.sphere_sums(xyz, r35, start, 1.0, -1.0).cpu().numpy()
    # union bounds: max over atoms <= union <= sum over atoms, for counts of all three classes
    for col in (0, 2, 4):
        a = per_atom[:, col].reshape(nres, 5)
        assert (per_res[:, col] <= a.sum(axis=1)).all() and (per_res[:, col] >= a.max(axis=1)).all()
    assert np.array_equal(per_res[:, 6], per_atom[:, 6].reshape(nres, 5).min(axis=1))     # valid iff every atom is valid
    assert np.array_equal(per_res[:, 7], per_atom[:, 7].reshape(nres, 5).sum(axis=1))     # candidates add up
    # singleton groups == ungrouped; a repeated atom changes nothing (set semantics); deterministic across runs
    single = dev.sphere_sums(xyz[:5000], r35[:5000], np.arange(5001, dtype=np.int32), 1.0, -1.0).cpu().numpy()
    assert np.array_equal(single[:, (0, 2, 4, 6, 7)], per_atom[:5000][:, (0, 2, 4, 6, 7)])
    gc.close(single[:, (1, 3, 5)], per_atom[:5000][:, (1, 3, 5)], rtol=1e-9, atol=1e-9)
    dup_xyz = np.concatenate((xyz[:500], xyz[:500])).reshape(2, 100, 5, 3).transpose(1, 0, 2, 3).reshape(-1, 3)
    dup = dev.sphere_sums(dup_xyz, np.full(len(dup_xyz), 3.5, np.float32), np.arange(0, 1001, 10, dtype=np.int32), 1.0, -1.0).cpu().numpy()
    assert np.array_equal(dup[:, :7], per_res[:100, :7])
    again = dev.sphere_sums(xyz, r35, start, 1.0, -1.0).cpu().numpy()
    assert np.array_equal(again, per_res)
    # the oracle on a random sample of the full-size problem
    g = orc.geom(hdr, hdr.origin, mv=_blas.probe())
    rho = vol.cpu().numpy()
    pick = rng.choice(nres, 40, replace=False)
    for k in pick:
        want = orc.sphere_union_sums(g, rho, xyz[5 * k:5 * k + 5], r35[:5], 1.0, -1.0)
        assert np.array_equal(per_res[k, (0, 2, 4, 6)], want[[0, 2, 4, 6]])
        gc.close(per_res[k, (1, 3, 5)], want[[1, 3, 5]], rtol=1e-9, atol=1e-9)
    atoms = rng.choice(len(xyz), 200, replace=False)
    want = orc.sphere_sums_batch(g, rho, xyz[atoms], r35[:200], 0.0)
    assert np.array_equal(per_atom[atoms, 0], want[:, 0])
    gc.close(per_atom[atoms, 1], want[:, 1], rtol=1e-9, atol=1e-9)


def test_blob_edge_cases():
    """Empty foreground, a single class, every voxel foreground (one giant blob: worst case for the union-find),
    capacity overflow with automatic retry, and a map whose stored grid is narrower than a vector load."""
    import torch
    from impl_cuda import CudaImpl
    from impl_oracle import OracleImpl
    from pdb_eda_b200 import ccp4, synthetic
    vol = synthetic.smoothNoiseMap(40, seed=5, sigma=1.0)
    dm = ccp4.parse(io.BytesIO(synthetic.ccp4Bytes(vol[:, :37, :35], (20.0,) * 3 + (90, 90, 90), (40, 40, 40))), "edge")
    cuda, orc = CudaImpl(dm), OracleImpl(dm)
    dev = dm.deviceMap
    none = dev.blob_label(50.0, -50.0)                               # nothing beyond the cutoffs
    assert all(p["n_voxels"] == 0 and p["n_blobs"] == 0 and len(p["crs"]) == 0 for p in none)
    only_green = dev.blob_label(1.5, 0.0)
    assert only_green[1] is None and only_green[0]["n_voxels"] > 0
    assert dm.createFullBlobList(-50.0) == []
    assert dev.blob_label(0.0, 0.0) == [None, None]
    # everything is foreground: one blob holding the whole unique volume
    lo = float(np.float32(vol.min())) - 1.0
    shifted = ccp4.parse(io.BytesIO(synthetic.ccp4Bytes(vol[:, :37, :35] - lo, (20.0,) * 3 + (90, 90, 90), (40, 40, 40))), "dense")
    full = shifted.deviceMap.blob_label(0.5, 0.0)[0]                  # all values >= 1 > 0.5
    assert full["n_voxels"] == 40 * 37 * 35 and full["n_blobs"] == 1 and int(full["label"].max()) == 0
    gc.close(full["stats"][0, 0].item(), 40 * 37 * 35)
    # capacity overflow -> the wrapper grows the buffers and retries; results equal the roomy run
    roomy = dev.blob_label(1.2, -1.2)
    tight = dev.blob_label(1.2, -1.2, cap_voxels=64, cap_blobs=8)
    for a, b in zip(roomy, tight):
        assert torch.equal(a["crs"], b["crs"]) and torch.equal(a["label"], b["label"]) and torch.equal(a["stats"], b["stats"])
    for (c1, l1, s1), (c2, l2, s2) in zip(cuda.full_blobs(1.2, -1.2), orc.full_blobs(1.2, -1.2)):
        assert np.array_equal(c1, c2) and np.array_equal(l1, l2)
        gc.close(s1, s2, rtol=1e-9, atol=1e-9)


def test_empty_inputs_everywhere():
    from pdb_eda_b200 import _device, ccp4, cutils, synthetic
    dm = ccp4.parse(io.BytesIO(synthetic.ccp4Bytes(np.zeros((8, 8, 8), np.float32), (4.0,) * 3 + (90, 90, 90), (8, 8, 8))), "z")
    dev = dm.deviceMap
    assert dm.meanDensity == 0.0 and dm.stdDensity == 0.0
    a, i, x = dev.symmetry_expand(np.zeros((0, 3)), np.zeros((2, 12)), np.zeros((27, 3)), [0, 0, 0], [1, 1, 1])
    assert len(a) == 0 and len(i) == 0 and x.shape == (0, 3)
    idx, dist = _device.nearest_atom(np.zeros((0, 3)), np.zeros((5, 3)))
    assert len(idx) == 0 and len(dist) == 0
    label, n = _device.cluster_crs(np.zeros((0, 3), np.int32))
    assert n == 0 and len(label) == 0
    assert len(cutils.overlapPairs(np.zeros((0, 3), np.int32), np.zeros(0, np.int32))) == 0
    assert cutils.crsStats(dm, np.zeros((0, 3), np.int32), None, None, 0).shape == (0, 8)
    assert cutils.createCrsLists([]) == [] and dm.createBlobList([]) == []
    assert cutils.getSphereCrsFromXyzList(dm, [], 1.0) == set() and cutils.testValidXyzList(dm, [], 1.0)
    assert cutils.sumOfAbs([], 0.5) == 0
    assert cutils.getSphereCrsFromXyz(dm, [1.0, 1.0, 1.0], 0.0) in ([], [(2, 2, 2)])
    with pytest.raises(Exception):
        _device.nearest_atom(np.zeros((2, 3)), np.zeros((0, 3)))      # np.argmin of an empty row raises in the reference too


@pytest.mark.parametrize("seed", range(12))
def test_fuzz_random_geometries(seed):
    """Random cells (orthogonal and skewed), axis orders, starts, partially stored / over-sampled maps: spheres, unions,
    per-atom clusters, blobs, symmetry images and point lookups against the oracle."""
    import itertools
    from impl_cuda import CudaImpl
    from impl_oracle import OracleImpl
    from pdb_eda_b200 import ccp4, synthetic
    rng = np.random.default_rng(3000 + seed)
    data, _ = cases.random_geometry(seed)
    dm = ccp4.parse(io.BytesIO(data), "fuzz%d" % seed)
    cuda, orc = CudaImpl(dm), OracleImpl(dm)
    atoms = cases.random_atoms(dm, 14, seed=seed + 50, margin=4.0).astype(np.float64)
    radii = rng.uniform(0.4, 2.6, len(atoms)).astype(np.float32)
    m, s = cuda.mean_std()
    cut = m + 1.1 * s
    for c in (0.0, cut, -cut):
        a, b = cuda.sphere_lists(atoms, radii, c), orc.sphere_lists(atoms, radii, c)
        assert np.array_equal(a[1], b[1]) and np.array_equal(a[0], b[0])
    gstart = np.array([0, 3, 3, 8, 14], dtype=np.int32)
    _cmp_sums(cuda.sphere_sums(atoms, radii, gstart, cut, -cut), orc.sphere_sums(atoms, radii, gstart, cut, -cut))
    _cmp_sums(cuda.sphere_sums(atoms, radii, None, cut, -cut), orc.sphere_sums(atoms, radii, None, cut, -cut))
    ca, cb = cuda.sphere_clouds(atoms, radii, cut), orc.sphere_clouds(atoms, radii, cut)
    assert np.array_equal(ca[0], cb[0]) and np.array_equal(ca[1], cb[1])
    gc.close(ca[2], cb[2])
    for (c1, l1, s1), (c2, l2, s2) in zip(cuda.full_blobs(m + 2.2 * s, -(m + 2.2 * s)), orc.full_blobs(m + 2.2 * s, -(m + 2.2 * s))):
        assert np.array_equal(c1, c2) and np.array_equal(l1, l2)
        gc.close(s1, s2, rtol=1e-9, atol=1e-9)
    crs = np.stack([rng.integers(-2 * dm.header.crsInterval[k], 3 * dm.header.crsInterval[k], 300) for k in range(3)], axis=1).astype(np.int32)
    va, oka = cuda.point_density(crs)
    vb, okb = orc.point_density(crs)
    assert np.array_equal(va, vb) and np.array_equal(oka, okb.astype(bool))
    assert np.array_equal(cuda.xyz2crs(atoms), orc.xyz2crs(atoms))
    ops = [np.concatenate((np.round(np.linalg.qr(rng.normal(size=(3, 3)))[0], 6), rng.uniform(-20, 20, (3, 1))), axis=1) for _ in range(3)]
    ops[0] = np.concatenate((np.eye(3), np.zeros((3, 1))), axis=1)
    shift = np.array([np.dot(dm.header.orthoMat, v) for v in itertools.product((-1, 0, 1), repeat=3)])
    lo, hi = atoms.min(axis=0) - 6, atoms.max(axis=0) + 6
    sa, sb = cuda.symmetry(atoms, ops, shift, lo, hi), orc.symmetry(atoms, ops, shift, lo, hi)
    assert np.array_equal(sa[0], sb[0]) and np.array_equal(sa[1], sb[1])
    gc.close(sa[2], sb[2], rtol=1e-12, atol=1e-11)


def test_voxel_pass_matches_the_individual_calls():
    """pipeline.VoxelPass (what bench.py times): device-resident step and the host-in / host-out call give the results
    of the separate entry points, and the blob part agrees with the oracle."""
    import torch
    from impl_oracle import OracleImpl
    from pdb_eda_b200 import ccp4, synthetic
    from pdb_eda_b200.pipeline import VoxelPass
    cell, n = (32.0, 32.0, 32.0, 90, 90, 90), (64, 64, 64)
    st = synthetic.polyAlaStructure(60, (0, 0, 0), cell[:3], seed=8)
    a, b = synthetic.mapPair(st, n, cell, seed=9)
    dens = ccp4.parse(io.BytesIO(synthetic.ccp4Bytes(a, cell, n)), "d")
    diff = ccp4.parse(io.BytesIO(synthetic.ccp4Bytes(b, cell, n)), "f")
    xyz = np.array([at.coord for at in st.get_atoms()], dtype=np.float64)
    radii = np.tile(np.array([0.78, 0.72, 0.66, 0.81, 0.84], dtype=np.float32), 60)
    start = np.arange(0, 301, 5, dtype=np.int32)
    vp = VoxelPass(dens.deviceMap, diff.deviceMap, xyz, radii, start, 3.5)
    vp.step()
    res = vp.results()
    cut = vp.density_cut
    want_cloud = dens.deviceMap.sphere_sums(xyz, radii, None, cut, 0.0).cpu().numpy()
    want_region = dens.deviceMap.sphere_sums(xyz, np.full(300, 3.5, np.float32), start, cut, 0.0).cpu().numpy()
    assert np.array_equal(res["cloud"], want_cloud) and np.array_equal(res["region"], want_region)
    orc = OracleImpl(diff)
    for tag, (crs, label, stats) in zip(("green", "red"), orc.full_blobs(vp.diff_cut, -vp.diff_cut)):
        key = (crs[:, 0].astype(np.int64) * 64 + crs[:, 1]) * 64 + crs[:, 2]
        assert np.array_equal(res[tag]["key"].astype(np.int64) & 0xFFFFFFFF, key) and np.array_equal(res[tag]["label"], label)
        gc.close(res[tag]["stats"], stats, rtol=1e-9, atol=1e-9)
    h_dens = torch.from_numpy(np.ascontiguousarray(a).reshape(-1)).pin_memory()
    h_diff = torch.from_numpy(np.ascontiguousarray(b).reshape(-1)).pin_memory()
    h_xyz = torch.from_numpy(xyz).pin_memory()
    vp.dens.rho.zero_()
    vp.diff.rho.zero_()                                             # the host call must bring the maps back itself
    again = vp.stepFromHost(h_dens, h_diff, h_xyz)
    assert np.array_equal(again["cloud"], want_cloud) and np.array_equal(again["region"], want_region)
    assert np.array_equal(again["green"]["label"], res["green"]["label"]) and np.array_equal(again["red"]["stats"], res["red"]["stats"])


@pytest.mark.parametrize("name", ["ortho", "perm"])
def test_xyz2crs_ties_and_near_ties(name):
    """xyz2crs rounds (x - origin) / gridLength half to even (pdb_eda/ccp4.py:297-299).  The device forms the quotient with a
    reciprocal and keeps the exact division for quotients within 1e-9 of a half-integer (round_quotient, pe_common.cuh): exact
    ties, their neighbours one and a few ulp away, and ordinary points must all land on the reference's index."""
    dm, cuda, orc = _impls(name)
    h = dm.header
    origin = np.asarray(h.origin, dtype=np.float64)
    gl = np.asarray(h.gridLength, dtype=np.float64)
    rng = np.random.default_rng(5)
    k = rng.integers(-40, 200, size=(400, 3)).astype(np.float64)
    pts = [origin + (k + 0.5) * gl]                                  # ties up to the rounding of the products
    for steps in (1, 2, 3, 17):
        pts.append(np.nextafter(pts[0], np.inf) if steps == 1 else pts[0] + steps * np.spacing(pts[0]))
        pts.append(np.nextafter(pts[0], -np.inf) if steps == 1 else pts[0] - steps * np.spacing(pts[0]))
    pts.append(origin + (k + rng.uniform(0.4999999, 0.5000001, size=k.shape)) * gl)
    pts.append(origin + rng.uniform(-40, 200, size=k.shape) * gl)
    xyz = np.concatenate(pts)
    got = cuda.xyz2crs(xyz)
    want = orc.xyz2crs(xyz)
    assert np.array_equal(got, want)
    # and against the arithmetic itself: Python's round() of the correctly rounded quotient
    q = (xyz - origin) / gl
    ref = np.rint(q).astype(np.int64)[:, [int(a) for a in h.map2crs]]
    assert np.array_equal(got, ref)
