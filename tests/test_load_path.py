"""The device load path of ``ccp4.parse`` (pdb_eda/ccp4.py:77-127): maps of 32 MB and more are streamed
from the file handle through a page-locked ring to HBM without a host copy; the host views are fetched back on demand."""
import io

import numpy as np
import pytest

from pdb_eda_b200 import ccp4, synthetic

pytestmark = pytest.mark.gpu

N = 208                                               # 208^3 x 4 B = 36 MB: above PINNED_MIN_BYTES, 4.3 ring slots
CELL = (N * 0.5,) * 3 + (90.0, 90.0, 90.0)


def _volume(seed):
    return np.random.default_rng(seed).standard_normal((N, N, N), dtype=np.float32)


def test_streamed_map_equals_the_file_payload():
    values = _volume(1)
    data = synthetic.ccp4Bytes(values, CELL, (N, N, N))
    assert len(data) - ccp4.HEADER_BYTES >= ccp4.PINNED_MIN_BYTES
    dm = ccp4.parse(io.BytesIO(data), "big")
    assert dm._host32 is None                         # streamed: nothing kept on the host
    flat = values.reshape(-1)
    assert abs(dm.meanDensity - flat.astype(np.float64).mean()) < 1e-9
    assert abs(dm.stdDensity - flat.astype(np.float64).std()) < 1e-9
    assert dm._host32 is None                         # statistics are device reductions
    assert np.array_equal(dm.densityArray, flat)      # fetched back bit for bit, file order
    assert dm.density.shape == (N, N, N) and dm.density[3, 2, 1] == float(values[3, 2, 1])
    assert dm.getPointDensityFromCrs([1, 2, 3]) == values[3, 2, 1]


def test_two_maps_in_a_row_reuse_the_ring():
    a, b = _volume(2), _volume(3)
    da = ccp4.parse(io.BytesIO(synthetic.ccp4Bytes(a, CELL, (N, N, N))), "a")
    db = ccp4.parse(io.BytesIO(synthetic.ccp4Bytes(b, CELL, (N, N, N))), "b")
    assert np.array_equal(db.densityArray, b.reshape(-1))
    assert np.array_equal(da.densityArray, a.reshape(-1))


def test_short_reads_are_accepted():
    """A handle whose readinto returns fewer bytes than asked for (sockets, pipes: ccp4.readFromURL)."""
    values = _volume(4)

    class Dribble(io.BytesIO):
        def readinto(self, buffer):
            return super().readinto(memoryview(buffer)[:1234567])

    dm = ccp4.parse(Dribble(synthetic.ccp4Bytes(values, CELL, (N, N, N))), "dribble")
    assert np.array_equal(dm.densityArray, values.reshape(-1))


def test_size_mismatch_is_refused_like_the_reference():
    data = synthetic.ccp4Bytes(_volume(5), CELL, (N, N, N))
    with pytest.raises(AssertionError):
        ccp4.parse(io.BytesIO(data[:-4]), "short")
    with pytest.raises(AssertionError):
        ccp4.parse(io.BytesIO(data + b"\0\0\0\0"), "long")


def test_writes_into_density_still_reach_the_device():
    """The reference's tests write into ``density`` (tests/test_ccp4.py:77-87); a streamed map has to honour that too."""
    values = _volume(6)
    dm = ccp4.parse(io.BytesIO(synthetic.ccp4Bytes(values, CELL, (N, N, N))), "w")
    before = dm.meanDensity
    dm.density[0, 0, 0] = 1000.0
    assert dm.getPointDensityFromCrs([0, 0, 0]) == 1000.0
    expected = values.reshape(-1).astype(np.float64)
    expected[0] = 1000.0
    assert abs(dm.meanDensity - expected.mean()) < 1e-9 and dm.meanDensity != before
