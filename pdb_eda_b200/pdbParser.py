"""PDB header reader -- host-side input of the symmetry expansion (mirror of ``pdb_eda.pdbParser``).

Only what the voxel path consumes matters here: the REMARK 290 SMTRY records become ``header.rotationMats``, a
list of 3x4 float64 operators in file order (pdb_eda/pdbParser.py:71-77), which
``DensityAnalysis._calculateSymmetryAtoms`` hands to ``createSymmetryAtoms``.  The remaining header fields
(resolution, R values, program, space group) are kept under the reference's attribute names
(pdb_eda/pdbParser.py:116-140) because the callers of the path print them.
"""
import re

import numpy as np

_HEADER_PATTERNS = (
    ("resolution", "REMARK   2 RESOLUTION", re.compile(r"RESOLUTION.(.+)ANGSTROMS"), False),
    ("rValue", "REMARK   3   R VALUE", re.compile(r"^REMARK   3   R VALUE            \(WORKING SET\) : (.+)$"), False),
    ("rFree", "REMARK   3   FREE R VALUE", re.compile(r"^REMARK   3   FREE R VALUE                     : (.+)$"), False),
    ("program", "REMARK   3   PROGRAM", re.compile(r"^REMARK   3   PROGRAM     : (.+)$"), True),
    ("spaceGroup", "REMARK 290 SYMMETRY OPERATORS FOR SPACE GROUP:",
     re.compile(r"^REMARK 290 SYMMETRY OPERATORS FOR SPACE GROUP: (.+)$"), True),
)
_SMTRY = re.compile(r"^REMARK 290   SMTRY(.+)$")
_ATOM_COLUMNS = (("recordType", 0, 6), ("serial", 6, 11), ("atomName", 12, 16), ("alternateLocation", 16, 17),
                 ("residueName", 17, 20), ("chainID", 21, 22), ("residueNumber", 22, 26), ("x", 30, 38), ("y", 38, 46),
                 ("z", 46, 54), ("occupancy", 54, 60), ("bFactor", 60, 66), ("element", 76, 78))


def readPDBfile(file):
    """PDBentry from a file name or an open text handle (pdb_eda/pdbParser.py:11-21)."""
    if isinstance(file, str):
        with open(file, "r") as handle:
            return parse(handle)
    return parse(file)


def parse(handle, mode="lite"):
    """PDBentry from a text handle; ``mode='lite'`` stops at the first ATOM record (pdb_eda/pdbParser.py:24-98)."""
    values = {"pdbid": 0, "date": 0, "method": 0, "resolution": 0, "rValue": 0, "rFree": 0, "program": 0, "spaceGroup": 0}
    rotationMats = []
    atoms = []
    models = 0
    for record in handle.readlines():
        if mode == "lite" and record.startswith("ATOM"):
            break
        if record.startswith("HEADER"):
            values["date"] = record[57:59].strip()
            values["pdbid"] = record[62:66].strip()
        elif record.startswith("EXPDTA"):
            values["method"] = record[6:36].strip().replace(" ", "_")
        elif record.startswith("MODEL"):
            models += 1
            if models > 1:
                break
        elif record.startswith("REMARK 290   SMTRY"):
            match = _SMTRY.search(record)
            if match:
                items = match.group(1).split()
                row, op = int(items[0]), int(items[1])
                if len(rotationMats) < op:
                    rotationMats.append(np.zeros((3, 4)))
                rotationMats[op - 1][row - 1] = [float(v) for v in items[2:6]]
        elif record.startswith("ATOM") or record.startswith("HETATM"):
            fields = {name: record[lo:hi].strip() for name, lo, hi in _ATOM_COLUMNS}
            fields["record"] = record.strip()
            atoms.append(Atom(fields))
        else:
            for key, prefix, pattern, underscore in _HEADER_PATTERNS:
                if record.startswith(prefix):
                    match = pattern.search(record)
                    if match:
                        text = match.group(1).strip()
                        values[key] = text.replace(" ", "_") if underscore else text
                    break
    header = PDBheader(values["pdbid"], values["date"], values["method"], values["resolution"], values["rValue"],
                       values["rFree"], values["program"], values["spaceGroup"], rotationMats)
    return PDBentry(header, atoms)


class PDBentry:
    def __init__(self, header, atoms):
        self.header = header
        self.atoms = atoms


class PDBheader:
    def __init__(self, PDBid, date, method, resolution, rValue, rFree, program, spaceGroup, rotationMats):
        self.pdbid = PDBid
        self.date = date
        self.method = method
        self.resolution = resolution
        self.rValue = rValue
        self.rFree = rFree
        self.program = program
        self.spaceGroup = spaceGroup
        self.rotationMats = rotationMats


class Atom:
    def __init__(self, keyValues):
        for key, value in keyValues.items():
            setattr(self, key, value)
