"""Start-up probe of the host BLAS's 3x3 mat-vec accumulation order.

The reference converts coordinates in skewed cells and builds symmetry images with ``np.dot`` on 3x3 matrices
(pdb_eda/ccp4.py:282, :300, :316; pdb_eda/cutils.pyx:98-99).  The order in which the three products of a row are
accumulated, and whether they are fused, is a property of the host's BLAS kernel, not of pdb_eda (SURVEY.md
App. A.13).  The CUDA kernels evaluate the same expression in a selectable order (``pe_geom.mv_perm`` /
``pe_geom.mv_fma``); this module finds the order that reproduces ``np.dot`` on this host.
"""
import itertools
from fractions import Fraction

import numpy as np

_cached = None


def _fma(a, b, c):
    """Correctly rounded a*b + c (Python 3.12 has no math.fma)."""
    return float(Fraction(a) * Fraction(b) + Fraction(c))


def _row(a, x, perm, fused):
    acc = a[perm[0]] * x[perm[0]]
    for k in perm[1:]:
        acc = _fma(a[k], x[k], acc) if fused else acc + a[k] * x[k]
    return acc


def probe(samples=64, seed=12345):
    """Returns ``(perm, fma)``: np.dot(M, x)[i] == accumulate(M[i, perm[0]]*x[perm[0]], perm[1], perm[2])."""
    global _cached
    if _cached is not None:
        return _cached
    rng = np.random.default_rng(seed)
    cases = []
    for _ in range(samples):
        m = rng.uniform(-100.0, 100.0, (3, 3))
        x = rng.uniform(-100.0, 100.0, 3)
        # the operand flavours the reference uses: list . list, ndarray . float32 vector
        cases.append((m, x, np.dot(m.tolist(), x.tolist())))
        x32 = x.astype(np.float32)
        cases.append((m, x32.astype(np.float64), np.dot(m, x32)))
    best = None
    for perm in itertools.permutations(range(3)):
        for fused in (0, 1):
            hits = 0
            total = 0
            for m, x, y in cases:
                for i in range(3):
                    total += 1
                    hits += _row([float(v) for v in m[i]], [float(v) for v in x], perm, fused) == float(y[i])
            if best is None or hits > best[0]:
                best = (hits, total, perm, fused)
            if hits == total:
                _cached = (tuple(perm), int(fused))
                return _cached
    # no candidate reproduces every row: keep the closest (voxels within 2 ulp of a sphere surface may differ)
    _cached = (tuple(best[2]), int(best[3]))
    return _cached
