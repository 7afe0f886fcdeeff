#!/usr/bin/env python3
"""Wall clock of the public DensityAnalysis calls on the C2 structure (bench.py's `e2e_api` key on its own, with a
cProfile of aggregateCloud).  usage: python profiles/api_c2.py [residues=8000] [n=384]"""
import cProfile
import os
import pstats
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402

res = int(sys.argv[1]) if len(sys.argv) > 1 else 8000
n = int(sys.argv[2]) if len(sys.argv) > 2 else 384
work = bench.build_workload(dict(n=n, cell=n * 0.5, residues=res), seed=2)
print(bench.api_timing(work))
print(bench.api_timing(work))
if os.environ.get("PROFILE"):
    import io
    from pdb_eda_b200 import densityAnalysis, structure, synthetic
    cell = work["cell"]
    b1 = synthetic.ccp4Bytes(work["fofc2"], cell, (n, n, n))
    b2 = synthetic.ccp4Bytes(work["fofc"], cell, (n, n, n))
    text = structure.formatPDB(work["structure"], remark290=synthetic.cartesianOperators("P 1", cell), cell=cell, spaceGroup="P 1")
    pr = cProfile.Profile()
    pr.enable()
    an = densityAnalysis.fromFile(io.StringIO(text), io.BytesIO(b1), io.BytesIO(b2))
    an.aggregateCloud()
    g = an.greenBlobList
    pr.disable()
    pstats.Stats(pr).sort_stats("cumulative").print_stats(25)
