#!/usr/bin/env python3
"""cProfile of the host side of one rank of multiple-structures mode (where the per-structure time goes)."""
import cProfile
import io
import os
import pstats
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pdb_eda_b200 import densityAnalysis, multi, structure, synthetic  # noqa: E402

densityAnalysis.setGlobals(synthetic.defaultParams())
items = []
for i in range(12):
    n = (64, 96, 128)[i % 3]
    cell = (n * 0.5,) * 3 + (90.0, 90.0, 90.0)
    st = synthetic.polyAlaStructure(max(30, n ** 3 // 1750), (0, 0, 0), cell[:3], seed=100 + i, residuesPerChain=200)
    a, b = synthetic.mapPair(st, (n, n, n), cell, seed=200 + i)
    text = structure.formatPDB(st, remark290=synthetic.cartesianOperators("P 21 21 21", cell), cell=cell, spaceGroup="P 21 21 21")
    items.append((text, synthetic.ccp4Bytes(a, cell, (n, n, n)), synthetic.ccp4Bytes(b, cell, (n, n, n))))


def loader(item):
    an = densityAnalysis.fromFile(io.StringIO(item[0]), io.BytesIO(item[1]), io.BytesIO(item[2]))
    _ = an.greenBlobList, an.redBlobList
    return an


multi.runMultipleStructures(items[:2], loader, None, None, "cuda")
pr = cProfile.Profile()
pr.enable()
multi.runMultipleStructures(items, loader, None, None, "cuda")
torch.cuda.synchronize()
pr.disable()
s = io.StringIO()
pstats.Stats(pr, stream=s).sort_stats("cumulative").print_stats(45)
print(s.getvalue()[:9000])
