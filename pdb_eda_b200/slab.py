"""Slab-decomposed blob labelling of one very large map across GPUs (BASELINE.json config 4: 1024^3 Fo-Fc).

The reference holds a whole map in one process and clusters with an N x N distance matrix
(pdb_eda/ccp4.py:123-124, :337-338; pdb_eda/cutils.pyx:55), so it cannot run such maps at all.  Here the section axis
(the slowest-varying axis in memory, pdb_eda/ccp4.py:338) is cut into one slab per rank:

  1. every rank labels its slab with the same fused threshold + CCL kernels as a whole map (``pe_blob_label``);
  2. ONE exchange step: the first-section plane of every slab (sparse: column, row, blob id) is all-gathered over
     NVLink and each rank pairs its last plane with its successor's first plane (``pe_overlap_pairs``:
     26-adjacency across the cut) -> equivalences between blob ids of neighbouring slabs;
  3. the (small) equivalence list and the per-blob smallest canonical keys are all-gathered; every rank runs the
     same min-id union and ranks the merged blobs by their smallest canonical (column-slowest) key, which is the
     reference's blob order (pdb_eda/cutils.pyx:59-69);
  4. per-blob sums (DensityBlob.fromCrsList, pdb_eda/ccp4.py:522-545) are all-reduced.

Steps 3-4 are device-agnostic torch code over a handful of collectives, so they are tested on CPU with gloo;
steps 1-2 are CUDA.  ``labelSlabsEmulated`` runs all "ranks" one after another on a single GPU (no kernel ever
waits on another), which is how the merge is verified against the whole-map labelling on one device.
"""
import ctypes

import numpy as np
import torch
import torch.distributed as dist

from . import _device
from . import cutils as utils
from ._lib import PeGeom


def slabRanges(nSections, world):
    """[s0, s1) of every rank: contiguous, as equal as possible."""
    base, extra = divmod(nSections, world)
    out, s = [], 0
    for r in range(world):
        n = base + (1 if r < extra else 0)
        out.append((s, s + n))
        s += n
    return out


def _slabGeom(full, s0, s1):
    """Geometry of the slab [s0, s1) as a map of its own: same columns / rows, its own section count."""
    g = PeGeom.from_buffer_copy(bytes(full))
    u2 = full.unique_ncrs[2]
    g.ncrs[2] = s1 - s0
    g.unique_ncrs[2] = max(min(s1, u2) - s0, 1)
    return g


class LocalSlab:
    """What one rank knows after labelling its slab: per sign (0 green, 1 red) the foreground voxels with global
    (c, r, s), value, local blob number, and the first-plane / last-plane voxels."""

    def __init__(self, fullGeom, rhoSlab, s0, s1, cutPos, cutNeg):
        self.s0, self.s1 = s0, s1
        self.U = (fullGeom.unique_ncrs[0], fullGeom.unique_ncrs[1], fullGeom.unique_ncrs[2])
        covered = min(s1, self.U[2]) - s0
        self.parts = []
        if covered <= 0:                       # the slab lies entirely in the repeated part of the map
            dev = rhoSlab.device
            for _ in range(2):
                self.parts.append({"crs": torch.zeros((0, 3), dtype=torch.int64, device=dev), "value": torch.zeros(0, device=dev),
                                   "label": torch.zeros(0, dtype=torch.int64, device=dev), "n_blobs": 0})
            return
        dm = _device.DeviceMap(_slabGeom(fullGeom, s0, s1), rhoSlab.reshape(-1))
        for part in dm.blob_label(cutPos, cutNeg):
            if part is None:
                dev = rhoSlab.device
                self.parts.append({"crs": torch.zeros((0, 3), dtype=torch.int64, device=dev), "value": torch.zeros(0, device=dev),
                                   "label": torch.zeros(0, dtype=torch.int64, device=dev), "n_blobs": 0})
                continue
            crs = part["crs"].long().clone()
            crs[:, 2] += s0
            self.parts.append({"crs": crs, "value": part["value"], "label": part["label"].long(), "n_blobs": part["n_blobs"]})

    def plane(self, k, section):
        """(column, row, local blob number) of the sign-k foreground voxels in global section ``section``."""
        p = self.parts[k]
        sel = p["crs"][:, 2] == section
        return torch.cat((p["crs"][sel][:, :2], p["label"][sel][:, None]), dim=1)

    def minKeys(self, k):
        """Smallest canonical key (c*U1 + r)*U2 + s of every local blob."""
        p = self.parts[k]
        key = (p["crs"][:, 0] * self.U[1] + p["crs"][:, 1]) * self.U[2] + p["crs"][:, 2]
        out = torch.full((p["n_blobs"],), torch.iinfo(torch.int64).max, dtype=torch.int64, device=key.device)
        if len(key):
            out.scatter_reduce_(0, p["label"], key, reduce="amin")
        return out


def boundaryPairs(lastPlane, firstPlaneNext, offsetHere, offsetNext):
    """Equivalences across one cut: blobs of this slab's last plane that are 26-adjacent to blobs of the next slab's
    first plane.  Planes are (column, row, local blob) lists; returns (m, 2) global blob ids.  CUDA (pe_overlap_pairs)."""
    if len(lastPlane) == 0 or len(firstPlaneNext) == 0:
        return torch.zeros((0, 2), dtype=torch.int64, device=lastPlane.device)
    crs = torch.cat((torch.cat((lastPlane[:, :2], torch.zeros_like(lastPlane[:, :1])), dim=1),
                     torch.cat((firstPlaneNext[:, :2], torch.ones_like(firstPlaneNext[:, :1])), dim=1))).to(torch.int32)
    owner = torch.cat((lastPlane[:, 2] + offsetHere, firstPlaneNext[:, 2] + offsetNext)).to(torch.int32)
    pairs = utils.overlapPairs(crs, owner, None, device=crs.device)
    return torch.from_numpy(pairs.astype(np.int64)).to(lastPlane.device)


def mergeBlobIds(nTotal, pairs, minKeys):
    """Min-id union of the blob ids joined by ``pairs`` and the reference's numbering of the merged blobs.
    Returns (new blob number of every old id, number of merged blobs).  Pure torch; runs on any device."""
    dev = minKeys.device
    root = torch.arange(nTotal, dtype=torch.int64, device=dev)
    if len(pairs):
        a, b = pairs[:, 0], pairs[:, 1]
        while True:
            m = torch.minimum(root[a], root[b])
            new = root.clone()
            new.scatter_reduce_(0, a, m, reduce="amin")
            new.scatter_reduce_(0, b, m, reduce="amin")
            new = new[new]                                   # pointer jumping
            if torch.equal(new, root):
                break
            root = new
    compKey = torch.full((nTotal,), torch.iinfo(torch.int64).max, dtype=torch.int64, device=dev)
    if nTotal:
        compKey.scatter_reduce_(0, root, minKeys, reduce="amin")
    isRoot = root == torch.arange(nTotal, device=dev)
    roots = torch.nonzero(isRoot)[:, 0]
    order = torch.argsort(compKey[roots], stable=True)       # blobs in the order of their smallest canonical key
    number = torch.empty(nTotal, dtype=torch.int64, device=dev)
    number[roots[order]] = torch.arange(len(roots), device=dev)
    return number[root], int(len(roots))


def _allGatherRagged(t, group):
    """all_gather of tensors whose first dimension differs per rank (pad to the maximum, trim after)."""
    world = dist.get_world_size(group)
    n = torch.tensor([t.shape[0]], dtype=torch.int64, device=t.device)
    sizes = [torch.zeros_like(n) for _ in range(world)]
    dist.all_gather(sizes, n, group=group)
    sizes = [int(s.item()) for s in sizes]
    cap = max(max(sizes), 1)
    pad = torch.zeros((cap,) + tuple(t.shape[1:]), dtype=t.dtype, device=t.device)
    pad[:t.shape[0]] = t
    out = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(out, pad, group=group)
    return [o[:s] for o, s in zip(out, sizes)]


def mergeDistributed(nLocalBlobs, firstPlane, lastPlane, minKeys, pairFn=boundaryPairs, group=None):
    """Steps 2-3 for one sign on this rank.  ``firstPlane`` / ``lastPlane``: (column, row, local blob) of this slab's
    boundary sections; ``minKeys``: smallest canonical key per local blob.  Returns (new number of every local blob,
    total number of merged blobs)."""
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    counts = _allGatherRagged(torch.tensor([[nLocalBlobs]], dtype=torch.int64, device=minKeys.device), group)
    counts = [int(c[0, 0].item()) for c in counts]
    offsets = np.concatenate(([0], np.cumsum(counts)))
    firsts = _allGatherRagged(firstPlane, group)             # the halo exchange: one boundary plane per slab
    if rank + 1 < world:
        pairs = pairFn(lastPlane, firsts[rank + 1], int(offsets[rank]), int(offsets[rank + 1]))
    else:
        pairs = torch.zeros((0, 2), dtype=torch.int64, device=minKeys.device)
    allPairs = torch.cat(_allGatherRagged(pairs, group))
    allKeys = torch.cat(_allGatherRagged(minKeys, group))
    number, nMerged = mergeBlobIds(int(offsets[-1]), allPairs, allKeys)
    return number[int(offsets[rank]):int(offsets[rank + 1])], nMerged


def _stats(fullDev, crs, value, label, nBlobs):
    """Per-blob n, sum rho, sum rho*xyz, sum xyz from global voxel coordinates (xyz by the pe_crs2xyz kernel)."""
    stats = torch.zeros((nBlobs, 8), dtype=torch.float64, device=crs.device)
    if len(crs):
        xyz = fullDev.crs2xyz(crs.to(torch.int32))
        d = value.double()
        cols = torch.cat((torch.ones_like(d)[:, None], d[:, None], d[:, None] * xyz, xyz), dim=1)
        stats.index_add_(0, label, cols)
    return stats


def labelSlabDistributed(fullHeader, rhoSlab, s0, s1, cutPos, cutNeg, origin=None, group=None):
    """One rank's call: label this rank's slab [s0, s1) of the map and merge across ranks.  Returns per sign a dict with
    this rank's voxels (global crs, value, global blob number) and the all-reduced per-blob statistics."""
    geom = _device.geom_from_header(fullHeader, origin)
    local = LocalSlab(geom, rhoSlab, s0, s1, cutPos, cutNeg)
    fullDev = _device.DeviceMap(_slabGeom(geom, s0, s1), rhoSlab.reshape(-1))
    fullDev.geom = geom                                         # coordinates use the whole map's geometry
    out = []
    for k in range(2):
        p = local.parts[k]
        number, nMerged = mergeDistributed(p["n_blobs"], local.plane(k, s0), local.plane(k, s1 - 1), local.minKeys(k), group=group)
        label = number[p["label"]] if len(p["label"]) else p["label"]
        stats = _stats(fullDev, p["crs"], p["value"], label, nMerged)
        dist.all_reduce(stats, op=dist.ReduceOp.SUM, group=group)
        out.append({"crs": p["crs"], "value": p["value"], "label": label, "n_blobs": nMerged, "stats": stats})
    return out


def labelSlabsEmulated(fullHeader, rho, world, cutPos, cutNeg, origin=None):
    """All ranks of ``labelSlabDistributed`` executed one after another on ONE GPU (``rho``: the whole map on the device).
    Returns per sign the complete voxel list in canonical order with global blob numbers and statistics."""
    geom = _device.geom_from_header(fullHeader, origin)
    ns, nr, nc = geom.ncrs[2], geom.ncrs[1], geom.ncrs[0]
    vol = rho.reshape(ns, nr, nc)
    ranges = slabRanges(ns, world)
    locals_ = [LocalSlab(geom, vol[s0:s1].contiguous(), s0, s1, cutPos, cutNeg) for s0, s1 in ranges]
    fullDev = _device.DeviceMap(geom, rho.reshape(-1))
    out = []
    for k in range(2):
        counts = [l.parts[k]["n_blobs"] for l in locals_]
        offsets = np.concatenate(([0], np.cumsum(counts)))
        pairs = [boundaryPairs(locals_[r].plane(k, ranges[r][1] - 1), locals_[r + 1].plane(k, ranges[r + 1][0]), int(offsets[r]),
                               int(offsets[r + 1])) for r in range(world - 1)]
        dev = rho.device
        allPairs = torch.cat(pairs) if pairs else torch.zeros((0, 2), dtype=torch.int64, device=dev)
        allKeys = torch.cat([l.minKeys(k) for l in locals_])
        number, nMerged = mergeBlobIds(int(offsets[-1]), allPairs, allKeys)
        crs = torch.cat([l.parts[k]["crs"] for l in locals_])
        value = torch.cat([l.parts[k]["value"] for l in locals_])
        label = torch.cat([number[l.parts[k]["label"] + int(offsets[r])] for r, l in enumerate(locals_)])
        U = locals_[0].U
        key = (crs[:, 0] * U[1] + crs[:, 1]) * U[2] + crs[:, 2]
        order = torch.argsort(key)
        crs, value, label = crs[order], value[order], label[order]
        out.append({"crs": crs, "value": value, "label": label, "n_blobs": nMerged,
                    "stats": _stats(fullDev, crs, value, label, nMerged), "n_pairs": int(len(allPairs))})
    return out
