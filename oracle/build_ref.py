#!/usr/bin/env python3
"""Build recipe for ``oracle/_ref`` -- the UNMODIFIED reference, installed for checking only.

TEST INFRASTRUCTURE.  Nothing under ``oracle/`` is part of the product; only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline / ``--impl reference`` legs may use it.

The reference (pdb_eda 2.7.1) is Python plus one Cython file (``pdb_eda/cutils.pyx``, built by
``setup.py:42`` with ``-O3``).  This script runs the reference's own ``setup.py`` through pip, from a
scratch copy (``/root/reference`` is read-only and the build writes next to the sources), and installs the
result into ``oracle/_ref`` (git-ignored, but shipped to the GPU box with the gpurun snapshot).  No reference
source file is copied into the tracked tree.

Run:  python oracle/build_ref.py            (no-op when /root/reference is absent, e.g. on the GPU box)
"""
import os
import shutil
import subprocess
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
REF_SRC = os.environ.get("PDB_EDA_REFERENCE", "/root/reference")
DEST = os.path.join(HERE, "_ref")


def have_ref():
    """True when a usable installed reference (with the compiled cutils) is present."""
    pkg = os.path.join(DEST, "pdb_eda")
    if not os.path.isdir(pkg):
        return False
    return any(f.startswith("cutils.") and f.endswith(".so") for f in os.listdir(pkg))


def build(force=False):
    if have_ref() and not force:
        return True
    if not os.path.isdir(REF_SRC):
        return False
    scratch = tempfile.mkdtemp(prefix="pdb_eda_src_")
    try:
        src = os.path.join(scratch, "src")
        shutil.copytree(REF_SRC, src)
        if os.path.isdir(DEST):
            shutil.rmtree(DEST)
        cmd = [sys.executable, "-m", "pip", "install", "--no-index", "--no-build-isolation", "--no-deps",
               "--find-links", "/opt/wheelhouse", "--target", DEST, src]
        subprocess.run(cmd, check=True, stdout=subprocess.DEVNULL)
    finally:
        shutil.rmtree(scratch, ignore_errors=True)
    return have_ref()


if __name__ == "__main__":
    ok = build(force="--force" in sys.argv)
    print("oracle/_ref:", "ready" if ok else "unavailable (no /root/reference here)")
