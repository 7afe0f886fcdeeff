#!/usr/bin/env python3
"""cProfile of the five public DensityAnalysis calls on the C2 structure (the `e2e_api` key of bench.py), one profile per call, on
the second (warm) pass.  Says where the host Python goes; the kernels behind these calls are ~1 ms in total.
usage: python profiles/api_profile.py [top]"""
import cProfile
import io
import os
import pstats
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import bench  # noqa: E402
from pdb_eda_b200 import densityAnalysis, structure, synthetic  # noqa: E402

top = int(sys.argv[1]) if len(sys.argv) > 1 else 18
work = bench.build_workload(bench.FULL, 0)
cell, n = work["cell"], work["n"]
densityAnalysis.setGlobals(synthetic.defaultParams())
b1 = synthetic.ccp4Bytes(work["fofc2"], cell, (n, n, n))
b2 = synthetic.ccp4Bytes(work["fofc"], cell, (n, n, n))
text = structure.formatPDB(work["structure"], remark290=synthetic.cartesianOperators("P 1", cell), cell=cell, spaceGroup="P 1")


def load():
    an = densityAnalysis.fromFile(io.StringIO(text), io.BytesIO(b1), io.BytesIO(b2))
    an.densityObj.meanDensity, an.diffDensityObj.meanDensity
    return an


def stage(label, fn, profile):
    torch.cuda.synchronize()
    pr = cProfile.Profile() if profile else None
    t = time.perf_counter()
    res = pr.runcall(fn) if profile else fn()
    torch.cuda.synchronize()
    dt = time.perf_counter() - t
    print("== %s: %.4f s%s" % (label, dt, " (under cProfile)" if profile else ""))
    if profile:
        s = io.StringIO()
        pstats.Stats(pr, stream=s).sort_stats("cumulative").print_stats(top)
        print("\n".join(line for line in s.getvalue().splitlines()[4:] if line.strip()))
    return res


for profile in (False, False, True):
    an = stage("load", load, profile)
    stage("aggregateCloud", an.aggregateCloud, profile)
    blobs = stage("green/red blob lists", lambda: (an.greenBlobList, an.redBlobList), profile)
    stage("blob statistics", lambda: an.calculateAtomSpecificBlobStatistics(blobs[0] + blobs[1]), profile)
    stage("residue region density", lambda: an.calculateResidueRegionDensity(bench.REGION_RADIUS), profile)
