#!/usr/bin/env python3
"""Small end-to-end pass over every kernel for compute-sanitizer (memcheck / racecheck): golden case 'tric' through the
CUDA adapter + a small DensityAnalysis run."""
import io
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np  # noqa: E402
import golden_checks as gc  # noqa: E402
from impl_cuda import CudaImpl  # noqa: E402
from pdb_eda_b200 import densityAnalysis, structure, synthetic  # noqa: E402

for name in ("tric", "over"):
    gold = gc.load(name)
    dm, _ = gc.header_and_bytes(gold)
    impl = CudaImpl(dm)
    gc.check_conversions(gold, impl)
    gc.check_points(gold, impl)
    gc.check_mean_std(gold, impl)
    gc.check_sum_abs(gold, impl)
    gc.check_sphere_lists(gold, impl)
    gc.check_sphere_sums(gold, impl)
    gc.check_sphere_unions(gold, impl)
    gc.check_clouds(gold, impl)
    gc.check_blobs(gold, impl)
    gc.check_cluster(gold, impl)
    gc.check_symmetry(gold, impl, dm)
    gc.check_nearest(gold, impl)
cell, n = (20.0, 20.0, 20.0, 90, 90, 90), (40, 40, 40)
st = synthetic.polyAlaStructure(25, (0, 0, 0), cell[:3], seed=3)
a, b = synthetic.mapPair(st, n, cell, seed=4)
text = structure.formatPDB(st, remark290=synthetic.cartesianOperators("P 21 21 21", cell), cell=cell, spaceGroup="P 21 21 21")
densityAnalysis.setGlobals(synthetic.defaultParams())
an = densityAnalysis.fromFile(io.StringIO(text), io.BytesIO(synthetic.ccp4Bytes(a, cell, n)), io.BytesIO(synthetic.ccp4Bytes(b, cell, n)))
an.aggregateCloud(minTotalElectrons=10.0)
an.calculateAtomSpecificBlobStatistics(an.greenBlobList + an.redBlobList)
an.calculateResidueRegionDiscrepancies(3.5)
an.residueMetrics()
print("SANITIZE_RUN_OK", an.densityElectronRatio)
