"""GPU tier: the drop-in claim of INTEGRATION.md section 1 -- the UNMODIFIED reference (oracle/_ref) with its operator
seam `utils` (pdb_eda/ccp4.py:16-19, pdb_eda/densityAnalysis.py:26-29) re-bound to pdb_eda_b200.cutils produces the
results of the reference running on its own Cython cutils.  Every call below goes through the reference's own
DensityMatrix / DensityBlob / DensityAnalysis code; only the thirteen seam functions run on the GPU."""
import io

import numpy as np
import pytest

import golden_checks as gc
from pdb_eda_b200 import structure, synthetic

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def setup(ref):
    ref_ccp4, ref_da, ref_cutils, ref_pp = ref
    cell, n = (24.0, 27.0, 30.0, 90, 90, 90), (48, 54, 60)
    st = synthetic.polyAlaStructure(30, (0, 0, 0), cell[:3], seed=5)
    a, b = synthetic.mapPair(st, n, cell, seed=6, crsStart=(3, -2, 5), axisOrder=(2, 1, 3))
    d1 = synthetic.ccp4Bytes(a, cell, n, crsStart=(3, -2, 5), axisOrder=(2, 1, 3))
    d2 = synthetic.ccp4Bytes(b, cell, n, crsStart=(3, -2, 5), axisOrder=(2, 1, 3))
    text = structure.formatPDB(st, remark290=synthetic.cartesianOperators("P 21 21 21", cell), cell=cell, spaceGroup="P 21 21 21")

    def make():
        dens, diff = ref_ccp4.parse(io.BytesIO(d1), "t"), ref_ccp4.parse(io.BytesIO(d2), "t")
        dens.densityCutoff = dens.meanDensity + 1.5 * dens.stdDensity
        diff.diffDensityCutoff = diff.meanDensity + 3 * diff.stdDensity
        return ref_da.DensityAnalysis("t", dens, diff, st, ref_pp.readPDBfile(io.StringIO(text)))
    return ref, make


def _blob_key(blobs):
    return [(sorted(b.crsList), b.totalDensity, b.volume, list(b.centroid)) for b in blobs]


def test_reference_code_on_the_gpu_seam(setup, monkeypatch):
    (ref_ccp4, ref_da, ref_cutils, _), make = setup
    from pdb_eda_b200 import cutils as gpu_utils
    want = make()
    atoms = list(want.biopdbObj.get_atoms())
    xyz = [a.coord for a in atoms[:6]]
    exp = dict(
        green=_blob_key(want.greenBlobList), red=_blob_key(want.redBlobList),
        clouds=[_blob_key(want.densityObj.findAberrantBlobs(a.coord, 0.8, want.densityObj.densityCutoff)) for a in atoms[:25]],
        sphere=want.densityObj.getSphereCrsFromXyz(atoms[0].coord, 1.7, 0), union=want.densityObj.findAberrantBlobs(xyz, 2.0, 0.1),
        total=want.densityObj.getTotalDensityFromXyz(atoms[1].coord, 1.5, want.densityObj.densityCutoff),
        point=[want.densityObj.getPointDensityFromXyz(a.coord) for a in atoms[:10]],
        absd=want.diffDensityObj.getTotalAbsDensity(want.diffDensityObj.diffDensityCutoff),
        sym=[(s.symmetry, np.asarray(s.coord, dtype=np.float64)) for s in want.symmetryAtoms],
        valid=[ref_cutils.testValidXyz(want.densityObj, np.asarray(a.coord) + 20.0, 1.5) for a in atoms[:5]],
    )
    want.aggregateCloud()
    # ---- re-bind the seam: the one-line change a maintainer makes
    monkeypatch.setattr(ref_ccp4, "utils", gpu_utils)
    monkeypatch.setattr(ref_da, "utils", gpu_utils)
    got = make()
    assert got.densityObj.createFullBlobList.__module__ == "pdb_eda.ccp4"     # still the reference's classes

    def same(a, b):
        assert len(a) == len(b)
        for (ca, ta, va, xa), (cb, tb, vb, xb) in zip(a, b):
            assert ca == cb
            gc.close([ta, va] + xa, [tb, vb] + xb, rtol=1e-9, atol=1e-9)
    same(_blob_key(got.greenBlobList), exp["green"])
    same(_blob_key(got.redBlobList), exp["red"])
    for a, e in zip(atoms[:25], exp["clouds"]):
        same(_blob_key(got.densityObj.findAberrantBlobs(a.coord, 0.8, got.densityObj.densityCutoff)), e)
    assert got.densityObj.getSphereCrsFromXyz(atoms[0].coord, 1.7, 0) == exp["sphere"]
    u = got.densityObj.findAberrantBlobs(xyz, 2.0, 0.1)
    assert sorted(sorted(b.crsList) for b in u) == sorted(sorted(b.crsList) for b in exp["union"])
    gc.close([got.densityObj.getTotalDensityFromXyz(atoms[1].coord, 1.5, got.densityObj.densityCutoff)], [exp["total"]])
    assert [float(got.densityObj.getPointDensityFromXyz(a.coord)) for a in atoms[:10]] == [float(v) for v in exp["point"]]
    gc.close([got.diffDensityObj.getTotalAbsDensity(got.diffDensityObj.diffDensityCutoff)], [exp["absd"]])
    gsym = got.symmetryAtoms
    assert [s.symmetry for s in gsym] == [s for s, _ in exp["sym"]]
    gc.close(np.array([np.asarray(s.coord, dtype=np.float64) for s in gsym]), np.array([c for _, c in exp["sym"]]), rtol=1e-12, atol=1e-10)
    assert [gpu_utils.testValidXyz(got.densityObj, np.asarray(a.coord) + 20.0, 1.5) for a in atoms[:5]] == exp["valid"]
    # the reference's own aggregateCloud (its Python loops, testOverlap / merge through the GPU seam)
    got.aggregateCloud()
    assert got.numVoxelsAggregated == want.numVoxelsAggregated
    gc.close([got.densityElectronRatio, got.totalAggregatedDensity], [want.densityElectronRatio, want.totalAggregatedDensity])
    assert dict(got.atomTypeOverlapCompleteness) == dict(want.atomTypeOverlapCompleteness)
