"""pdb_eda_b200 -- the voxel hot path of pdb_eda (MoseleyBioinformaticsLab/pdb_eda 2.7.1) on NVIDIA B200.

Host code is Python (CCP4 header / orthogonalisation parsing, PDB side); voxels live in HBM and every voxel loop
is a hand-written sm_100a CUDA kernel behind the C ABI of ``include/pdbeda_b200.h``.  The public names follow the
reference: :mod:`pdb_eda_b200.ccp4`, :mod:`pdb_eda_b200.cutils` (drop-in for ``pdb_eda.cutils``),
:mod:`pdb_eda_b200.densityAnalysis` (``fromPDBid``, ``fromFile``, ``DensityAnalysis``), :mod:`pdb_eda_b200.pdbParser`.
"""
__version__ = "0.1.0"


def __getattr__(name):
    # lazy: importing the package must not need torch / the CUDA library (the CPU test tier imports pieces of it)
    if name in ("fromPDBid", "fromFile", "DensityAnalysis"):
        from . import densityAnalysis
        return getattr(densityAnalysis, name)
    raise AttributeError(name)
