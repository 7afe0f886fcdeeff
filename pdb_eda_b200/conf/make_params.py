#!/usr/bin/env python3
"""Regenerates ``pdb_eda_b200/conf/optimized_params.json`` from the reference's parameter table.

The optimised parameter set (427 RES_ATOM names -> 100 atom types; per-type radii and b-factor slopes, bonded atoms,
electron counts, leaving atoms) is DATA the reference ships inside its package and loads at import
(pdb_eda/conf/optimized_params.json, pdb_eda/densityAnalysis.py:32-46).  The voxel path needs it for every real entry
(the radii feed the sphere kernels), so the table is vendored as package data -- re-encoded by this script, values
untouched -- instead of being looked up in an installed pdb_eda.

usage: python pdb_eda_b200/conf/make_params.py [/root/reference/pdb_eda/conf/optimized_params.json]
"""
import json
import os
import sys

KEYS = ("radii", "slopes", "bonded_atoms", "full_atom_name_map_electrons", "full_atom_name_map_atom_type", "leaving_atoms")


def main():
    src = sys.argv[1] if len(sys.argv) > 1 else "/root/reference/pdb_eda/conf/optimized_params.json"
    with open(src) as fh:
        params = json.load(fh)
    missing = [k for k in KEYS if k not in params]
    if missing:
        raise SystemExit("missing keys in %s: %s" % (src, missing))
    out = os.path.join(os.path.dirname(os.path.abspath(__file__)), "optimized_params.json")
    with open(out, "w") as fh:
        json.dump({k: params[k] for k in KEYS}, fh, sort_keys=True, separators=(",", ":"))
        fh.write("\n")
    print("%s: %d atom names, %d atom types" % (out, len(params["full_atom_name_map_atom_type"]), len(params["radii"])))


if __name__ == "__main__":
    main()
