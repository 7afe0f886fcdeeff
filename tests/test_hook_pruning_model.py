"""CPU model of the neighbour selection of the sparse blob stage's hooking phase (pe_blob.cu, P2), checked against
scipy.ndimage.label.

A voxel (c, r, s) hooks only "predecessor" neighbours: those in the columns (c-1, r-1), (c-1, r), (c-1, r+1), (c, r-1)
at sections s-1, s, s+1 (its own column is chained by runs of consecutive sections).  The kernel drops hooks that are
implied by others: a neighbour at the voxel's own section stands in for every other predecessor neighbour within one
row and one column of it, and, when no neighbour shares the section, the same per side (s-1 / s+1).  This test runs
exactly that selection in Python with a plain union-find and asserts that the components equal the 26-connected
components -- the induction argument in the kernel comment, executed on dense random masks where it matters most.
"""
import numpy as np
import pytest
import scipy.ndimage as ndi

COLS = ((-1, -1), (-1, 0), (-1, 1), (0, -1))  # (dc, dr) of the four predecessor columns, in the kernel's order


def _find(parent, x):
    while parent[x] != x:
        parent[x] = parent[parent[x]]
        x = parent[x]
    return x


def _union(parent, a, b):
    a, b = _find(parent, a), _find(parent, b)
    if a != b:
        parent[max(a, b)] = min(a, b)


def _label_with_pruned_hooks(mask, prune):
    nc, nr, ns = mask.shape                      # [c][r][s]: the reference's scan order is C order of this array
    ident = -np.ones(mask.shape, dtype=np.int64)
    ident[mask] = np.arange(int(mask.sum()))
    parent = list(range(int(mask.sum())))
    hooks = 0
    for c, r, s in zip(*np.nonzero(mask)):
        me = ident[c, r, s]
        if s > 0 and mask[c, r, s - 1]:          # runs along the section axis are chained at initialisation
            _union(parent, me, ident[c, r, s - 1])
        sel = []
        for dc, dr in COLS:                      # per column: same section, else s-1 and s+1
            cc, rr = c + dc, r + dr
            m = {}
            if 0 <= cc < nc and 0 <= rr < nr:
                if mask[cc, rr, s]:
                    m[0] = ident[cc, rr, s]
                else:
                    if s > 0 and mask[cc, rr, s - 1]:
                        m[-1] = ident[cc, rr, s - 1]
                    if s + 1 < ns and mask[cc, rr, s + 1]:
                        m[1] = ident[cc, rr, s + 1]
            sel.append(m)
        if prune:
            if 0 in sel[1]:
                sel[0], sel[2], sel[3] = {}, {}, {}
            elif 0 in sel[3]:
                sel[0], sel[1] = {}, {}
            elif 0 in sel[0]:
                sel[1], sel[3] = {}, {}
            elif 0 in sel[2]:
                sel[1] = {}
            else:
                for side in (-1, 1):
                    if side in sel[1]:
                        for j in (0, 2, 3):
                            sel[j].pop(side, None)
                    elif side in sel[3]:
                        sel[0].pop(side, None)
        for m in sel:
            for other in m.values():
                _union(parent, me, other)
                hooks += 1
    roots = np.array([_find(parent, i) for i in range(len(parent))])
    return roots, hooks


@pytest.mark.parametrize("density", [0.08, 0.3, 0.55, 0.8])
def test_pruned_hooks_give_the_26_connected_components(density):
    rng = np.random.default_rng(int(density * 100))
    for shape in ((9, 8, 10), (6, 12, 7)):
        mask = rng.random(shape) < density
        lab, n = ndi.label(mask, structure=np.ones((3, 3, 3), dtype=bool))
        flat = lab[mask]
        full, hooks_full = _label_with_pruned_hooks(mask, prune=False)
        pruned, hooks_pruned = _label_with_pruned_hooks(mask, prune=True)
        for roots in (full, pruned):
            assert len(np.unique(roots)) == n
            # same partition: one root per scipy label and vice versa
            assert len(set(zip(roots.tolist(), flat.tolist()))) == n
            # a root is the first voxel of its blob in scan order (canonical numbering)
            first = {}
            for i, l in enumerate(flat.tolist()):
                first.setdefault(l, i)
            assert all(roots[i] == first[l] for i, l in enumerate(flat.tolist()))
        assert hooks_pruned <= hooks_full
