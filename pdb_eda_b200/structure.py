"""Minimal structure hierarchy with the Biopython surface the voxel path touches.

The reference walks a ``Bio.PDB`` structure (pdb_eda/densityAnalysis.py:596-643, :905, :937, :971, :1023-1033;
SURVEY.md App. B.3): ``get_residues()``, ``get_atoms()``, ``header['resolution']``, residue ``id`` /
``resname`` / ``child_list`` / ``parent``, atom ``coord`` (float32[3]) / ``name`` / ``element`` /
``get_occupancy()`` / ``get_bfactor()`` / ``parent``.  When Biopython is installed the loaders use it; otherwise
(this image has no Biopython) they fall back on this small PDB-format reader, which yields the same surface.
Any object with that surface works with :class:`pdb_eda_b200.densityAnalysis.DensityAnalysis`.
"""
import itertools

import numpy as np

from ._gc import pausedGC

_ATOM_RECORDS = ("ATOM  ", "HETATM")
_FLOAT32 = np.dtype(np.float32)                       # native float32 arrays carry this very object as their dtype


class Atom:
    def __init__(self, name, coord, bfactor, occupancy, altloc=" ", fullname=None, serial_number=0, element="", parent=None):
        self.name = name
        self.fullname = fullname if fullname is not None else name
        self.coord = coord if (type(coord) is np.ndarray and coord.dtype is _FLOAT32) else np.asarray(coord, dtype=np.float32)
        self.bfactor = float(bfactor)
        self.occupancy = float(occupancy)
        self.altloc = altloc
        self.serial_number = serial_number
        self.element = element
        self.parent = parent

    def get_occupancy(self):
        return self.occupancy

    def get_bfactor(self):
        return self.bfactor

    def get_name(self):
        return self.name

    def get_coord(self):
        return self.coord

    def get_parent(self):
        return self.parent

    def __repr__(self):
        return "<Atom %s>" % self.name


class _Entity:
    def __init__(self, id):
        self.id = id
        self.parent = None
        self.child_list = []

    def add(self, child):
        child.parent = self
        self.child_list.append(child)
        return child

    def get_parent(self):
        return self.parent

    def __iter__(self):
        return iter(self.child_list)

    def __len__(self):
        return len(self.child_list)


class Residue(_Entity):
    def __init__(self, id, resname, segid=" "):
        super().__init__(id)
        self.resname = resname
        self.segid = segid

    def get_atoms(self):
        return iter(self.child_list)

    def get_resname(self):
        return self.resname


class Chain(_Entity):
    def get_residues(self):
        return iter(self.child_list)

    def get_atoms(self):
        for residue in self.child_list:
            yield from residue.child_list


class Model(_Entity):
    def get_chains(self):
        return iter(self.child_list)

    def get_residues(self):
        for chain in self.child_list:
            yield from chain.child_list

    def get_atoms(self):
        for residue in self.get_residues():
            yield from residue.child_list


class Structure(_Entity):
    def __init__(self, id):
        super().__init__(id)
        self.header = {"resolution": None}

    def get_models(self):
        return iter(self.child_list)

    def get_chains(self):
        for model in self.child_list:
            yield from model.child_list

    def get_residues(self):
        for chain in self.get_chains():
            yield from chain.child_list

    def get_atoms(self):
        for residue in self.get_residues():
            yield from residue.child_list


def _parsePDBColumns(lines, structureId):
    """Fast path of ``parsePDB`` for the common case -- one model, no alternate locations: the fixed-width columns of all
    ATOM / HETATM records are converted with numpy in one go (coordinates, occupancies, b-factors, residue numbers, names);
    the Python loop that remains only creates the objects.  Returns None when the file needs the general reader."""
    atomLines = [line for line in lines if line.startswith(_ATOM_RECORDS)]
    resolution = None
    for line in (lines if len(atomLines) == len(lines) else [line for line in lines if not line.startswith(_ATOM_RECORDS)]):
        rec = line[0:6]
        if rec == "MODEL " or rec == "ENDMDL":
            return None
        if rec == "REMARK" and line.startswith("REMARK   2 RESOLUTION."):
            try:
                resolution = float(line[23:30])
            except ValueError:
                pass
    structure = Structure(structureId)
    structure.header["resolution"] = resolution
    n = len(atomLines)
    if n == 0:
        return structure
    try:
        width = len(atomLines[0])
        if width >= 79 and set(map(len, atomLines)) == {width}:                 # the usual case: every record is 80 columns wide
            raw = np.frombuffer("".join(atomLines).encode("ascii"), dtype="S1").reshape(n, width)
            if width < 80:
                raw = np.concatenate((raw, np.full((n, 80 - width), b" ", dtype="S1")), axis=1)
        else:
            raw = np.frombuffer("".join(line.rstrip("\n").ljust(80)[:80] for line in atomLines).encode("ascii"), dtype="S1").reshape(n, 80)
    except UnicodeEncodeError:
        return None
    col = lambda a, b: np.ascontiguousarray(raw[:, a:b]).view("S%d" % (b - a)).ravel()

    def text(a, b, fn=lambda t: t):
        """Column a:b as a list of str, fn applied once per DISTINCT value (atom names, residue names ... repeat)."""
        values, inverse = np.unique(col(a, b), return_inverse=True)
        table = np.empty(len(values), dtype=object)
        table[:] = [fn(v.decode("ascii")) for v in values.tolist()]
        return table[np.asarray(inverse).reshape(-1)]

    if (col(16, 17) != b" ").any():
        return None                                   # alternate locations: the general reader picks the highest occupancy
    try:
        coords = np.stack([col(30, 38).astype(np.float64), col(38, 46).astype(np.float64), col(46, 54).astype(np.float64)], axis=1).astype(np.float32)
        number = lambda a, b, blank: np.where(col(a, b) == b" " * (b - a), blank.rjust(b - a), col(a, b))
        occ = number(54, 60, b"1.0").astype(np.float64).tolist()
        bfac = number(60, 66, b"0.0").astype(np.float64).tolist()
        serial = number(6, 11, b"0").astype(np.int64).tolist()
        resseq = col(22, 26).astype(np.int64)
    except ValueError:
        return None
    fullnames = text(12, 16).tolist()
    nameValues, nameInverse = np.unique(col(12, 16), return_inverse=True)
    strippedTable = [v.decode("ascii").strip() for v in nameValues.tolist()]
    idOfName = {}
    nameIds = np.array([idOfName.setdefault(name, len(idOfName)) for name in strippedTable], dtype=np.int64)[np.asarray(nameInverse).reshape(-1)]
    nameTable = np.empty(len(strippedTable), dtype=object)
    nameTable[:] = strippedTable
    names = nameTable[np.asarray(nameInverse).reshape(-1)].tolist()
    resnames = text(17, 20, str.strip)
    chainIds = text(21, 22)
    icodes = text(26, 27)
    elements = text(76, 78, lambda t: t.strip().upper()).tolist()
    hetatm = col(0, 6) == b"HETATM"
    hetflags = np.full(n, " ", dtype=object)
    if hetatm.any():
        hetflags[hetatm] = [("W" if r in ("HOH", "WAT") else "H_" + r) for r in resnames[hetatm].tolist()]
    # a new residue starts where (chain, hetero flag, number, insertion code) changes; a key that comes back later in the file
    # (interleaved residues) needs the general reader's dictionary
    change = np.ones(n, dtype=bool)
    change[1:] = (chainIds[1:] != chainIds[:-1]) | (hetflags[1:] != hetflags[:-1]) | (resseq[1:] != resseq[:-1]) | (icodes[1:] != icodes[:-1])
    starts = np.flatnonzero(change)
    keys = set()
    for k in starts.tolist():
        key = (chainIds[k], hetflags[k], int(resseq[k]), icodes[k])
        if key in keys:
            return None
        keys.add(key)
    # a repeated atom name inside a residue: general reader
    if len(np.unique((np.cumsum(change) - 1) * len(idOfName) + nameIds)) < n:
        return None
    model = structure.add(Model(0))
    chains = {}
    resnameList, chainList, hetList, icodeList, resseqList = resnames.tolist(), chainIds.tolist(), hetflags.tolist(), icodes.tolist(), resseq.tolist()
    bounds = starts.tolist() + [n]
    residues = []
    for a in bounds[:-1]:
        chain = chains.get(chainList[a])
        if chain is None:
            chain = chains[chainList[a]] = model.add(Chain(chainList[a]))
        residues.append(chain.add(Residue((hetList[a], resseqList[a], icodeList[a]), resnameList[a])))
    # the atoms in one C-level sweep (map over the prepared columns), their residue handed to the constructor
    owners = np.empty(len(residues), dtype=object)
    owners[:] = residues
    owners = np.repeat(owners, np.diff(bounds)).tolist()
    atoms = list(map(Atom, names, list(coords), bfac, occ, itertools.repeat(" "), fullnames, serial, elements, owners))
    for residue, a, b in zip(residues, bounds[:-1], bounds[1:]):
        residue.child_list = atoms[a:b]
    return structure


@pausedGC
def parsePDB(handle, structureId="xxxx"):
    """Reads ATOM / HETATM / MODEL records of a PDB-format text handle (or file name) into a Structure.

    Alternate locations: the highest-occupancy one is kept (Biopython's DisorderedAtom default selection).
    """
    if isinstance(handle, str):
        with open(handle, "r") as fh:
            return parsePDB(fh, structureId)
    lines = handle.readlines() if hasattr(handle, "readlines") else list(handle)
    fast = _parsePDBColumns(lines, structureId)
    if fast is not None:
        return fast
    handle = lines
    structure = Structure(structureId)
    model = None
    chains = {}
    residues = {}
    atoms = {}
    model_serial = 0
    for line in handle:
        rec = line[0:6]
        if rec == "MODEL ":
            model = structure.add(Model(model_serial))
            model_serial += 1
            chains, residues, atoms = {}, {}, {}
        elif rec == "ENDMDL":
            model = None
        elif rec in ("ATOM  ", "HETATM"):
            if model is None:
                model = structure.add(Model(model_serial))
                model_serial += 1
                chains, residues, atoms = {}, {}, {}
            fullname = line[12:16]
            name = fullname.strip()
            altloc = line[16]
            resname = line[17:20].strip()
            chain_id = line[21]
            resseq = int(line[22:26])
            icode = line[26]
            if rec == "HETATM":
                hetflag = "W" if resname in ("HOH", "WAT") else "H_" + resname
            else:
                hetflag = " "
            coord = (float(line[30:38]), float(line[38:46]), float(line[46:54]))
            occupancy = float(line[54:60]) if line[54:60].strip() else 1.0
            bfactor = float(line[60:66]) if line[60:66].strip() else 0.0
            element = line[76:78].strip().upper() if len(line) >= 78 else ""
            serial = int(line[6:11]) if line[6:11].strip() else 0
            chain = chains.get(chain_id)
            if chain is None:
                chain = chains[chain_id] = model.add(Chain(chain_id))
            rkey = (chain_id, hetflag, resseq, icode)
            residue = residues.get(rkey)
            if residue is None:
                residue = residues[rkey] = chain.add(Residue((hetflag, resseq, icode), resname))
            atom = Atom(name, coord, bfactor, occupancy, altloc, fullname, serial, element)
            akey = rkey + (name,)
            prev = atoms.get(akey)
            if prev is None:
                atoms[akey] = residue.add(atom)
            elif altloc != " " and occupancy > prev.occupancy:
                atom.parent = residue
                residue.child_list[residue.child_list.index(prev)] = atom
                atoms[akey] = atom
        elif line.startswith("REMARK   2 RESOLUTION."):
            try:
                structure.header["resolution"] = float(line[23:30])
            except ValueError:
                pass
    return structure


def formatPDB(structure, remark290=None, cell=None, spaceGroup="P 1", resolution=2.0):
    """PDB-format text of a structure (header with REMARK 290 operators, CRYST1, ATOM records)."""
    out = ["HEADER    SYNTHETIC STRUCTURE                     01-JAN-00   XXXX              \n",
           "REMARK   2 RESOLUTION.   %5.2f ANGSTROMS.                                       \n" % resolution]
    if remark290 is not None:
        out.append("REMARK 290 SYMMETRY OPERATORS FOR SPACE GROUP: %s\n" % spaceGroup)
        for k, mat in enumerate(remark290):
            for row in range(3):
                out.append("REMARK 290   SMTRY%d %3d %9.6f %9.6f %9.6f %14.5f\n"
                           % (row + 1, k + 1, mat[row][0], mat[row][1], mat[row][2], mat[row][3]))
    if cell is not None:
        out.append("CRYST1%9.3f%9.3f%9.3f%7.2f%7.2f%7.2f %-11s%4d\n" % (tuple(cell) + (spaceGroup, 1)))
    serial = 1
    for chain in structure.get_chains():
        for residue in chain:
            rec = "ATOM  " if residue.id[0] == " " else "HETATM"
            for atom in residue:
                nm = atom.name
                field = (" " + nm).ljust(4) if len(nm) < 4 else nm
                out.append("%s%5d %s%s%3s %s%4d%s   %8.3f%8.3f%8.3f%6.2f%6.2f          %2s\n"
                           % (rec, serial, field, atom.altloc, residue.resname, chain.id, residue.id[1], residue.id[2],
                              float(atom.coord[0]), float(atom.coord[1]), float(atom.coord[2]), atom.occupancy, atom.bfactor,
                              atom.element.rjust(2)))
                serial += 1
    out.append("END\n")
    return "".join(out)
