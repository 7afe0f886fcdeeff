"""CPU tier: the column-wise fast path of the PDB reader (structure._parsePDBColumns) yields exactly what the general
record-by-record reader yields, and steps aside for files it does not cover (alternate locations, several models,
repeated residue keys, short records)."""
import io

import numpy as np

from pdb_eda_b200 import structure, synthetic


def _general(text):
    saved = structure._parsePDBColumns
    structure._parsePDBColumns = lambda *a: None
    try:
        return structure.parsePDB(io.StringIO(text))
    finally:
        structure._parsePDBColumns = saved


def _same(a, b):
    fa, fb = list(a.get_atoms()), list(b.get_atoms())
    assert len(fa) == len(fb) and a.header == b.header
    for x, y in zip(fa, fb):
        assert (x.name, x.fullname, x.bfactor, x.occupancy, x.element, x.serial_number, x.altloc) == \
               (y.name, y.fullname, y.bfactor, y.occupancy, y.element, y.serial_number, y.altloc)
        assert np.array_equal(x.coord, y.coord) and x.coord.dtype == y.coord.dtype == np.float32
        assert x.parent.id == y.parent.id and x.parent.resname == y.parent.resname and x.parent.parent.id == y.parent.parent.id
        assert type(x.name) is str and type(x.parent.id[1]) is int and type(x.parent.resname) is str
    assert [len(r) for r in a.get_residues()] == [len(r) for r in b.get_residues()]


def test_fast_reader_equals_general_reader():
    st = synthetic.polyAlaStructure(300, (0, 0, 0), (60.0,) * 3, seed=4, residuesPerChain=70, hetero=9)
    text = structure.formatPDB(st, remark290=synthetic.cartesianOperators("P 21 21 21", (60.0, 60.0, 60.0, 90, 90, 90)),
                               cell=(60.0, 60.0, 60.0, 90, 90, 90), spaceGroup="P 21 21 21", resolution=1.85)
    fast = structure._parsePDBColumns(text.splitlines(True), "x")
    assert fast is not None and fast.header["resolution"] == 1.85
    _same(fast, _general(text))
    st.header["resolution"] = 1.85          # the generated structure itself, for the round trip through formatPDB
    for x, y in zip(structure.parsePDB(io.StringIO(text)).get_atoms(), st.get_atoms()):
        assert x.name == y.name and np.array_equal(x.coord, y.coord) and x.parent.id == y.parent.id


def test_fast_reader_steps_aside():
    st = synthetic.polyAlaStructure(12, (0, 0, 0), (30.0,) * 3, seed=5)
    text = structure.formatPDB(st, cell=(30.0, 30.0, 30.0, 90, 90, 90))
    lines = text.splitlines(True)
    first = next(i for i, l in enumerate(lines) if l.startswith("ATOM"))
    # blank occupancy / b-factor columns take the defaults in both readers
    blank = lines[:]
    blank[first] = blank[first][:54] + " " * 12 + blank[first][66:]
    _same(structure.parsePDB(io.StringIO("".join(blank))), _general("".join(blank)))
    assert list(structure.parsePDB(io.StringIO("".join(blank))).get_atoms())[0].occupancy == 1.0
    # alternate locations: general reader (the highest occupancy wins)
    alt = lines[:]
    alt[first] = alt[first][:16] + "A" + alt[first][17:54] + "  0.40" + alt[first][60:]
    alt.insert(first + 1, alt[first][:16] + "B" + alt[first][17:30] + "   1.000   2.000   3.000" + "  0.60" + alt[first][60:])
    assert structure._parsePDBColumns(alt, "x") is None
    kept = list(structure.parsePDB(io.StringIO("".join(alt))).get_atoms())[0]
    assert kept.altloc == "B" and kept.occupancy == 0.6
    # several models, short records
    assert structure._parsePDBColumns(["MODEL        1\n"] + lines, "x") is None
    short = [l[:60].rstrip() + "\n" if l.startswith("ATOM") else l for l in lines]
    _same(structure.parsePDB(io.StringIO("".join(short))), _general("".join(short)))


def test_repeated_atom_name_goes_to_the_general_reader():
    """A residue that lists an atom name twice (no altloc flag) is the general reader's business: it keeps ONE atom per name."""
    st = synthetic.polyAlaStructure(6, (0, 0, 0), (30.0,) * 3, seed=6)
    lines = structure.formatPDB(st, cell=(30.0, 30.0, 30.0, 90, 90, 90)).splitlines(True)
    first = next(i for i, l in enumerate(lines) if l.startswith("ATOM"))
    twice = lines[:first + 1] + [lines[first]] + lines[first + 1:]
    assert structure._parsePDBColumns(twice, "x") is None
    assert structure._parsePDBColumns(lines, "x") is not None
    # the same NAME in two different residues is the usual case and stays on the fast path
    _same(structure.parsePDB(io.StringIO("".join(lines))), _general("".join(lines)))
    # names that differ only in their padding ("CA  " / " CA ") are one name to Biopython
    padded = lines[:first + 1] + [lines[first][:12] + (lines[first][13:16] + " " if lines[first][12] == " " else " " + lines[first][12:15])
                                  + lines[first][16:]] + lines[first + 1:]
    if padded[first + 1][12:16].strip() == lines[first][12:16].strip() and padded[first + 1][12:16] != lines[first][12:16]:
        assert structure._parsePDBColumns(padded, "x") is None


def test_paused_gc_restores_the_collector():
    import gc
    from pdb_eda_b200._gc import pausedGC

    @pausedGC
    def inner():
        assert not gc.isenabled()
        return 7

    @pausedGC
    def outer(fail):
        assert not gc.isenabled()
        assert inner() == 7 and not gc.isenabled()     # nested: the inner call leaves the outer pause alone
        if fail:
            raise ValueError("x")
        return 1

    assert gc.isenabled()
    assert outer(False) == 1 and gc.isenabled()
    try:
        outer(True)
    except ValueError:
        pass
    assert gc.isenabled()                               # switched back on when the call raises
    gc.disable()
    try:
        assert outer(False) == 1 and not gc.isenabled()  # a caller that runs without the collector keeps it off
    finally:
        gc.enable()
