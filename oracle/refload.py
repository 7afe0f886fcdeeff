"""Import the installed reference from ``oracle/_ref`` (TEST INFRASTRUCTURE -- never used by the product).

Biopython is not in this image; the reference only needs ``import Bio.PDB`` to succeed at import time
(``pdb_eda/densityAnalysis.py:18``) and then works on any duck-typed structure object,
so an empty stand-in module is registered before importing it.
"""
import os
import sys
import types

HERE = os.path.dirname(os.path.abspath(__file__))
REF_DIR = os.path.join(HERE, "_ref")


def available():
    pkg = os.path.join(REF_DIR, "pdb_eda")
    return os.path.isdir(pkg) and any(f.startswith("cutils.") and f.endswith(".so") for f in os.listdir(pkg))


def load():
    """Returns (ccp4, densityAnalysis, cutils, pdbParser) modules of the real reference."""
    if not available():
        raise RuntimeError("oracle/_ref is not built; run `python oracle/build_ref.py` where /root/reference exists")
    if "Bio" not in sys.modules:
        try:
            import Bio.PDB  # noqa: F401
        except Exception:
            bio = types.ModuleType("Bio")
            biopdb = types.ModuleType("Bio.PDB")
            bio.PDB = biopdb
            sys.modules["Bio"] = bio
            sys.modules["Bio.PDB"] = biopdb
    if REF_DIR not in sys.path:
        sys.path.insert(0, REF_DIR)
    import pdb_eda.ccp4 as ref_ccp4
    import pdb_eda.densityAnalysis as ref_da
    import pdb_eda.cutils as ref_cutils
    import pdb_eda.pdbParser as ref_pdbparser
    assert ref_ccp4.utils.__name__ == "pdb_eda.cutils", "reference fell back to pure-Python utils"
    return ref_ccp4, ref_da, ref_cutils, ref_pdbparser
