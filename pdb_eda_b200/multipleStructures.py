"""``multiple`` mode command line: many PDB entries, sharded over the GPUs of the node (pdb_eda/multipleStructures.py:61-194).

Usage:
    python -m pdb_eda_b200 multiple <pdbid-file> <out-result-file> [--out-format json|csv] [--params FILE] [--data-dir DIR] [--workers W]
    torchrun --nproc-per-node N -m pdb_eda_b200 multiple ...        (one rank per GPU; rank 0 writes the result)

<pdbid-file>: JSON list or whitespace-separated text of PDB ids.  --data-dir: analyse DIR/<id>.ccp4, DIR/<id>_diff.ccp4,
DIR/pdb<id>.ent(.gz) instead of downloading.  Result: per-entry ``stats`` and ``diffs`` as in ``analyzePDBID``
(pdb_eda/multipleStructures.py:320-356), plus the all-reduced cumulative statistics.
"""
import argparse
import csv
import functools
import json
import os
import sys

import torch
import torch.distributed as dist

from . import densityAnalysis, multi

statsHeaders = ['density_electron_ratio', 'voxel_volume', 'num_voxels_aggregated', 'total_aggregated_electrons', 'num_atoms_analyzed',
                'num_residue_clouds_analyzed', 'num_domain_clouds_analyzed', 'atom_overlap_completeness']


def readIds(path):
    text = open(path).read()
    try:
        ids = json.loads(text)
    except ValueError:
        ids = text.split()
    return [str(i).lower() for i in ids]


def loadEntry(dataDir, pdbid):
    """DensityAnalysis of one entry: downloaded (``dataDir`` empty) or read from DIR/<id>.ccp4, DIR/<id>_diff.ccp4, DIR/pdb<id>.ent(.gz)."""
    if not dataDir:
        return densityAnalysis.fromPDBid(pdbid)
    pdb = os.path.join(dataDir, "pdb" + pdbid + ".ent.gz")
    if not os.path.isfile(pdb):
        pdb = os.path.join(dataDir, "pdb" + pdbid + ".ent")
    an = densityAnalysis.fromFile(pdb, os.path.join(dataDir, pdbid + ".ccp4"), os.path.join(dataDir, pdbid + "_diff.ccp4"))
    if an:
        an.pdbid = pdbid
    return an


def makeLoader(dataDir):
    return functools.partial(loadEntry, dataDir)     # picklable: usable by the host worker pool


def main(argv=None):
    p = argparse.ArgumentParser(prog="pdb_eda_b200 multiple")
    p.add_argument("pdbid_file")
    p.add_argument("out_result_file")
    p.add_argument("--out-format", default="json", choices=["json", "csv"])
    p.add_argument("--params", default="")
    p.add_argument("--data-dir", default="")
    p.add_argument("--workers", type=int, default=1, help="host worker processes per rank (all drive the rank's GPU)")
    args = p.parse_args(argv)
    if args.params:
        densityAnalysis.setGlobals(json.load(open(args.params)))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if world > 1 and not dist.is_initialized():
        torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", "0")))
        dist.init_process_group("nccl")
    rank = dist.get_rank() if dist.is_initialized() else 0
    ids = readIds(args.pdbid_file)
    types = sorted(densityAnalysis.paramsGlobal["radii"])
    summary = multi.runMultipleStructures(ids, makeLoader(args.data_dir), None, types, workers=args.workers)
    if rank == 0:
        cols = summary["columns"]
        full = {}
        for row in summary["rows"]:
            rec = dict(zip(cols, row.tolist()))
            pdbid = ids[int(rec["index"])]
            full[pdbid] = {"pdbid": pdbid, "stats": {h: rec[h] for h in statsHeaders},
                           "diffs": {t: rec["diff:" + t] for t in types}, "execution_time": rec["execution_time"]}
        with open(args.out_result_file, "w", newline="") if args.out_result_file != "-" else sys.stdout as out:
            if args.out_format == "csv":
                writer = csv.writer(out)
                writer.writerow(["pdbid"] + statsHeaders + types)
                for rec in full.values():
                    writer.writerow([rec["pdbid"]] + [rec["stats"][h] for h in statsHeaders] + [rec["diffs"][t] for t in types])
            else:
                print(json.dumps({"entries": full, "cumulative": summary["cumulative"]}, indent=2, sort_keys=True), file=out)
    if dist.is_initialized():
        dist.barrier()


if __name__ == "__main__":
    main()
