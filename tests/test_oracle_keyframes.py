"""CPU: oracle/keyframe_oracle.py (KeyframeExtractor.GenerateFromST3DForIntervals + compute_overlapping_CC_groups restated) against
the outputs of the unmodified reference on four seeded videos x five frame intervals (tests/golden/keyframes.npz): key-frame
pixels and the sorted (time, bounding box) lists, identical."""
import numpy as np
import pytest

from oracle import keyframe_oracle as KO
from oracle.gen_golden_grouping import RUNS
from oracle.gen_golden_keyframes import unpack_st3d


def expected(z, name):
    n, h, w = (int(v) for v in z[name + "/shape"])
    content = np.unpackbits(z[name + "/keyframes"], axis=-1)[:, :, :w].astype(bool)
    times = {}
    for row in z[name + "/times"]:
        times.setdefault(int(row[0]), []).append(tuple(row[1:]))
    segs = [tuple(int(v) for v in s) for s in z[name + "/segments"]]
    return n, h, w, segs, content, [times.get(s, []) for s in range(len(segs))]


def check(keyframes, times, content, ref_times):
    assert len(keyframes) == len(content)
    for k, c in zip(keyframes, content):
        assert k.dtype == np.uint8 and k.shape == c.shape + (3,)
        np.testing.assert_array_equal(k[:, :, 0] == 0, c)
        assert set(np.unique(k)) <= {0, 255} and (k[:, :, 0] == k[:, :, 1]).all() and (k[:, :, 0] == k[:, :, 2]).all()
    for got, ref in zip(times, ref_times):
        assert [tuple(float(v) for v in row) for row in got] == ref


@pytest.mark.parametrize("name", sorted(RUNS))
def test_keyframe_oracle_equals_reference(golden, name):
    z = golden("keyframes.npz")
    n, h, w, segs, content, ref_times = expected(z, name)
    ages, images, bounds = unpack_st3d(z, name + "/")
    kfs, times = KO.keyframes_for_intervals([40.0 * t for t in range(n)], h, w, ages, images, bounds, segs)
    check(kfs, times, content, ref_times)
