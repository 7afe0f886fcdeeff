"""CPU tier: the restatement oracle (oracle/pe_oracle.c) against the golden vectors of the real reference."""
import pytest

import golden_checks as gc
from impl_oracle import OracleImpl


@pytest.fixture(scope="module", params=gc.CASES)
def case(request):
    gold = gc.load(request.param)
    dm, _ = gc.header_and_bytes(gold)
    return gold, dm, OracleImpl(dm)


def test_header(case):
    gold, dm, _ = case
    gc.check_header(gold, dm.header)


def test_conversions(case):
    gc.check_conversions(case[0], case[2])


def test_points(case):
    gc.check_points(case[0], case[2])


def test_mean_std_sum_abs(case):
    gc.check_mean_std(case[0], case[2])
    gc.check_sum_abs(case[0], case[2])


def test_sphere_lists(case):
    gc.check_sphere_lists(case[0], case[2])


def test_sphere_sums(case):
    gc.check_sphere_sums(case[0], case[2])


def test_sphere_unions(case):
    gc.check_sphere_unions(case[0], case[2])


def test_clouds(case):
    gc.check_clouds(case[0], case[2])


def test_blobs(case):
    gc.check_blobs(case[0], case[2])


def test_cluster(case):
    gc.check_cluster(case[0], case[2])


def test_symmetry_and_nearest(case):
    gold, dm, impl = case
    gc.check_symmetry(gold, impl, dm)
    gc.check_nearest(gold, impl)


def test_em_origin_is_flagged():
    """A header with non-zero words 47-49 takes the reference's originEM branch (pdb_eda/ccp4.py:281-284), on which its sphere
    enumeration degenerates; the loader says so instead of reproducing the bug silently."""
    import struct
    import warnings
    from pdb_eda_b200 import ccp4, synthetic
    raw = bytearray(synthetic.ccp4Header((8, 8, 8), (4.0, 4.0, 4.0, 90, 90, 90), (8, 8, 8)))
    struct.pack_into("<3f", raw, 4 * 46, 1.0, 1.0, 1.0)        # futureUse[-3:]
    struct.pack_into("<3f", raw, 4 * 49, 2.0, 3.0, 4.0)        # originEM
    with warnings.catch_warnings(record=True) as caught:
        warnings.simplefilter("always")
        hdr = ccp4.DensityHeader.fromFileHeader(bytes(raw))
    assert isinstance(hdr.origin, list) and hdr.origin == [2.0, 3.0, 4.0]
    assert any("EM origin" in str(w.message) for w in caught)
