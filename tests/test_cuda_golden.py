"""GPU tier: the CUDA library (through the C ABI) against the golden vectors of the real reference."""
import pytest

import golden_checks as gc

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module", params=gc.CASES)
def case(request):
    from impl_cuda import CudaImpl
    gold = gc.load(request.param)
    dm, _ = gc.header_and_bytes(gold)
    return gold, dm, CudaImpl(dm)


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
