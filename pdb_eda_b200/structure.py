"""Minimal structure hierarchy with the Biopython surface the voxel path touches.

The reference walks a ``Bio.PDB`` structure (pdb_eda/densityAnalysis.py:596-643, :905, :937, :971, :1023-1033;
SURVEY.md App. B.3): ``get_residues()``, ``get_atoms()``, ``header['resolution']``, residue ``id`` /
``resname`` / ``child_list`` / ``parent``, atom ``coord`` (float32[3]) / ``name`` / ``element`` /
``get_occupancy()`` / ``get_bfactor()`` / ``parent``.  When Biopython is installed the loaders use it; otherwise
(this image has no Biopython) they fall back on this small PDB-format reader, which yields the same surface.
Any object with that surface works with :class:`pdb_eda_b200.densityAnalysis.DensityAnalysis`.
"""
import numpy as np


class Atom:
    def __init__(self, name, coord, bfactor, occupancy, altloc=" ", fullname=None, serial_number=0, element=""):
        self.name = name
        self.fullname = fullname if fullname is not None else name
        self.coord = np.asarray(coord, dtype=np.float32)
        self.bfactor = float(bfactor)
        self.occupancy = float(occupancy)
        self.altloc = altloc
        self.serial_number = serial_number
        self.element = element
        self.parent = None

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


def parsePDB(handle, structureId="xxxx"):
    """Reads ATOM / HETATM / MODEL records of a PDB-format text handle (or file name) into a Structure.

    Alternate locations: the highest-occupancy one is kept (Biopython's DisorderedAtom default selection).
    """
    if isinstance(handle, str):
        with open(handle, "r") as fh:
            return parsePDB(fh, structureId)
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
