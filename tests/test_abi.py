"""CPU tier: the C-ABI shared library loads and exports every symbol include/pdbeda_b200.h declares.
No compute call is made here (there is no GPU in this tier); argument validation runs on the host."""
import ctypes
import os
import re

import pytest

from pdb_eda_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "pdbeda_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(pe_[a-z0-9_]+)\s*\(", text)))


def test_every_declared_symbol_is_exported_and_bound():
    lib = _lib.load()
    names = declared_symbols()
    assert len(names) >= 25
    for name in names:
        assert hasattr(lib, name), "libpdbeda_b200.so does not export %s" % name
        assert name in _lib.SIGNATURES, "no ctypes prototype for %s" % name
    for name in _lib.SIGNATURES:
        assert name in names, "%s is bound but not declared in the header" % name


def test_abi_version_and_struct_layout():
    lib = _lib.load()
    assert lib.pe_abi_version() == 1
    # struct pe_geom: 7*3 + 1 + 3 + 1 int32 = 26 ints = 104 bytes, then 3 + 3 + 9 + 9 doubles
    assert ctypes.sizeof(_lib.PeGeom) == 104 + 24 * 8
    assert _lib.PeGeom.grid_length.offset == 104


def test_argument_errors_are_reported_without_a_device():
    lib = _lib.load()
    rc = lib.pe_map_mean_std(None, 0, None, None, None)
    assert rc == -1
    assert b"pe_map_mean_std" in lib.pe_last_error()
    g = _lib.PeGeom()
    rc = lib.pe_sphere_sums(ctypes.byref(g), None, 1, None, None, 1, None, 0.0, 0.0, None, None, None)
    assert rc == -1 and b"ncrs" in lib.pe_last_error()
    with pytest.raises(_lib.PdbEdaLibError):
        _lib.check(rc, "pe_sphere_sums")


def test_no_cpu_fallback():
    """Without a CUDA device the product raises instead of computing on the host."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    import numpy as np
    from pdb_eda_b200 import cutils
    with pytest.raises(_lib.PdbEdaLibError):
        cutils.createCrsLists([(0, 0, 0), (1, 1, 1)])
    with pytest.raises(_lib.PdbEdaLibError):
        cutils.sumOfAbs(np.ones(4, dtype=np.float32), 0.5)


def test_product_does_not_import_the_oracle():
    pkg = os.path.join(ROOT, "pdb_eda_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in text and "from oracle" not in text and "pe_oracle" not in text, f
