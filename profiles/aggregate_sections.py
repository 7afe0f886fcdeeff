import io, os, sys, time, re
sys.path.insert(0, '/root/repo')
import numpy as np, torch
from pdb_eda_b200 import densityAnalysis, structure, synthetic
src = open('/root/repo/pdb_eda_b200/densityAnalysis.py').read()
# instrument: insert checkpoints at '# ----' comment lines inside aggregateCloud
lines = src.split('\n')
out = []; inside = False; k = 0
for ln in lines:
    if ln.startswith('    def aggregateCloud'): inside = True
    elif inside and ln.startswith('    def ') or ln.startswith('    @staticmethod'): inside = False
    if inside and ln.strip().startswith('# ----'):
        ind = ln[:len(ln) - len(ln.lstrip())]
        out.append(ind + "torch.cuda.synchronize(); _T.append((%d, __import__('time').perf_counter()))" % k); k += 1
    out.append(ln)
code = '\n'.join(out).replace('from . import', 'from pdb_eda_b200 import')
ns = {'__name__': 'pdb_eda_b200.densityAnalysis_prof', '_T': []}
code = code.replace('class DensityAnalysis(object):', '_T = []\nclass DensityAnalysis(object):')
import types
mod = types.ModuleType('pdb_eda_b200.densityAnalysis_prof'); mod.__package__ = 'pdb_eda_b200'
exec(compile(code, 'da_prof', 'exec'), mod.__dict__)
mod.setGlobals(synthetic.defaultParams())
n = 128; cell = (64.0,) * 3 + (90.0,) * 3
st = synthetic.polyAlaStructure(1200, (0, 0, 0), cell[:3], seed=5, residuesPerChain=200)
a, b = synthetic.mapPair(st, (n, n, n), cell, seed=6)
text = structure.formatPDB(st, remark290=synthetic.cartesianOperators("P 21 21 21", cell), cell=cell, spaceGroup="P 21 21 21")
for rep in range(3):
    an = mod.fromFile(io.StringIO(text), io.BytesIO(synthetic.ccp4Bytes(a, cell, (n, n, n))), io.BytesIO(synthetic.ccp4Bytes(b, cell, (n, n, n))))
    torch.cuda.synchronize(); mod._T.clear(); t0 = time.perf_counter()
    an.aggregateCloud(); torch.cuda.synchronize(); t1 = time.perf_counter()
T = mod._T
print('total %.1f ms' % ((t1 - t0) * 1e3))
prev = t0
for k, t in T:
    print('  before section %d: +%.2f ms' % (k, (t - prev) * 1e3)); prev = t
print('  tail: +%.2f ms' % ((t1 - prev) * 1e3))
