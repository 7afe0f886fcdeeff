#!/usr/bin/env python3
"""BASELINE.json config 1 head to head: the UNMODIFIED reference (oracle/_ref, Cython cutils, one host core) and this package on
the same synthetic 96^3 P2(1)2(1)2(1) map pair with a ~2.5k-atom structure -- wall-clock per public API call and result parity.
usage: python profiles/c1_head_to_head.py [residues=500]"""
import io
import os
import sys
import time
import warnings

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import refload  # noqa: E402  (measurement script: the reference is the thing being timed, not a fallback)
from pdb_eda_b200 import densityAnalysis, structure, synthetic  # noqa: E402

warnings.simplefilter("ignore")
nres = int(sys.argv[1]) if len(sys.argv) > 1 else 500
ref_ccp4, ref_da, ref_cutils, ref_pp = refload.load()
n, cell = (96, 96, 96), (48.0, 48.0, 48.0, 90, 90, 90)
st = synthetic.polyAlaStructure(nres, (0, 0, 0), cell[:3], seed=1, residuesPerChain=250)
d1, d2 = synthetic.mapPair(st, n, cell, seed=7)
b1, b2 = synthetic.ccp4Bytes(d1, cell, n), synthetic.ccp4Bytes(d2, cell, n)
text = structure.formatPDB(st, remark290=synthetic.cartesianOperators("P 21 21 21", cell), cell=cell, spaceGroup="P 21 21 21")
densityAnalysis.setGlobals(ref_da.paramsGlobal)
mask = {"ALA": ["N", "CA", "C"]}


def timed(fn):
    if torch.cuda.is_available():
        torch.cuda.synchronize()
    t = time.perf_counter()
    out = fn()
    if torch.cuda.is_available():
        torch.cuda.synchronize()
    return out, time.perf_counter() - t


def load_ref():
    dens, diff = ref_ccp4.parse(io.BytesIO(b1), "c1"), ref_ccp4.parse(io.BytesIO(b2), "c1")
    dens.densityCutoff = dens.meanDensity + 1.5 * dens.stdDensity
    diff.diffDensityCutoff = diff.meanDensity + 3 * diff.stdDensity
    return ref_da.DensityAnalysis("c1", dens, diff, st, ref_pp.readPDBfile(io.StringIO(text)))


def load_mine():
    return densityAnalysis.fromFile(io.StringIO(text), io.BytesIO(b1), io.BytesIO(b2))


load_mine().aggregateCloud()                       # CUDA context, BLAS probe, library load
steps = [
    ("load (parse maps + structure, mean/std)", None),
    ("aggregateCloud", lambda a: a.aggregateCloud()),
    ("greenBlobList + redBlobList", lambda a: (a.greenBlobList, a.redBlobList)),
    ("symmetryAtoms", lambda a: a.symmetryAtoms),
    ("calculateAtomSpecificBlobStatistics", lambda a: a.calculateAtomSpecificBlobStatistics(a.greenBlobList + a.redBlobList)),
    ("calculateResidueRegionDensity (3.5 A, atom mask)", lambda a: a.calculateResidueRegionDensity(3.5, 1.5, "", mask)),
    ("calculateResidueRegionDiscrepancies (3.5 A, mask)", lambda a: a.calculateResidueRegionDiscrepancies(3.5, 3.0, "ALA", mask)),
]
r, tr0 = timed(load_ref)
m, tm0 = timed(load_mine)
m.densityObj.densityCutoff, m.diffDensityObj.diffDensityCutoff = r.densityObj.densityCutoff, r.diffDensityObj.diffDensityCutoff
rows = [(steps[0][0], tr0, tm0, "-")]
for label, fn in steps[1:]:
    ro, tr = timed(lambda: fn(r))
    mo, tm = timed(lambda: fn(m))
    same = "-"
    if label == "aggregateCloud":
        same = (r.numVoxelsAggregated == m.numVoxelsAggregated and abs(r.densityElectronRatio / m.densityElectronRatio - 1) < 1e-9 and
                len(r.residueCloudDescriptions) == len(m.residueCloudDescriptions) and len(r.domainCloudDescriptions) == len(m.domainCloudDescriptions))
    elif label.startswith("green"):
        same = all(len(x) == len(y) and all(p.crsList == q.crsList for p, q in zip(x, y)) for x, y in zip(ro, mo))
    elif label == "symmetryAtoms":
        same = len(ro) == len(mo) and np.allclose(np.array([np.asarray(a.coord, float) for a in ro]), np.array([np.asarray(a.coord, float) for a in mo]), rtol=1e-12, atol=1e-10)
    elif label.startswith("calculateAtomSpecific"):
        same = len(ro) == len(mo) and all(x[3] == y[3] and x[5:10] == y[5:10] and abs(x[0] - y[0]) <= 1e-9 * max(1.0, abs(x[0])) for x, y in zip(ro, mo))
    else:
        same = len(ro) == len(mo) and all(x[:4] == y[:4] and np.allclose(np.array(x[4:], float), np.array(y[4:], float), rtol=1e-9, atol=1e-9) for x, y in zip(ro, mo))
    rows.append((label, tr, tm, same))
print("C1 head to head: 96^3 P2(1)2(1)2(1), %d atoms; reference = unmodified pdb_eda 2.7.1 (Cython cutils) on one host core, %s" % (
    5 * nres, torch.cuda.get_device_name(0)))
print("%-52s %12s %12s %9s  %s" % ("API call", "reference s", "this repo s", "speed-up", "results match"))
for label, tr, tm, same in rows:
    print("%-52s %12.3f %12.4f %8.0fx  %s" % (label, tr, tm, tr / tm, same))
print("%-52s %12.3f %12.4f %8.0fx" % ("total", sum(x[1] for x in rows), sum(x[2] for x in rows), sum(x[1] for x in rows) / sum(x[2] for x in rows)))
