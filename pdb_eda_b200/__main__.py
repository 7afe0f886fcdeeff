"""``python -m pdb_eda_b200 <mode> ...`` -- mode dispatch (pdb_eda/__main__.py:29-66).  Modes on the voxel path: single, multiple."""
import sys


def main():
    if len(sys.argv) < 2 or sys.argv[1] in ("-h", "--help"):
        print(__doc__)
        print("usage: python -m pdb_eda_b200 (single | multiple) ...")
        return 0
    mode, argv = sys.argv[1], sys.argv[2:]
    if mode == "single":
        from . import singleStructure
        return singleStructure.main(argv)
    if mode == "multiple":
        from . import multipleStructures
        return multipleStructures.main(argv)
    print("unknown mode %r: the voxel path offers 'single' and 'multiple' (contacts / generate / optimize are out of scope)" % mode, file=sys.stderr)
    return 2


if __name__ == "__main__":
    sys.exit(main())
