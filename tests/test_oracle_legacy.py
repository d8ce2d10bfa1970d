"""CPU: the NumPy restatement of the four legacy accessmath_lib exports (oracle/legacy_oracle.py) against the golden
vectors captured from the reference's own compiled C (tests/golden/legacy_ops.npz, oracle/gen_golden_legacy.py) and,
when oracle/_ref is present, against that library directly on fresh random inputs.  Bar: bit-exact (uint8 and fp64)."""
import numpy as np
import pytest

from oracle import legacy_oracle as L
from oracle.gen_golden_legacy import legacy_inputs

INPUTS = legacy_inputs()


def run_oracle(name, d):
    if name.startswith("ahe"):
        return L.adapthisteq(d["gray"], d["slope"], d["gx"], d["gy"])
    if name.startswith("comb"):
        return L.combine_results(d["board"], d["eq"], d["thr"])
    t, b, a, dv = L.speaker_detection(d["frame"], d["last"], d["thr"], d["jump"])
    return np.concatenate([b, a, dv, [float(t)]])


@pytest.mark.parametrize("name", sorted(INPUTS))
def test_oracle_equals_reference_golden(golden, name):
    z = golden("legacy_ops.npz")
    got = run_oracle(name, INPUTS[name])
    assert got.dtype == z[name].dtype
    np.testing.assert_array_equal(got, z[name])
    if name.startswith("ahe"):
        d = INPUTS[name]
        h, w = d["gray"].shape
        cdf = L.region_cdf(d["gray"], w // 5, w - 3, h // 4, h - 2, d["slope"])
        np.testing.assert_array_equal(cdf, z[name + "_cdf"])


def test_oracle_equals_compiled_reference_random():
    lib = L.ref()
    if lib is None:
        pytest.skip("oracle/_ref not built")
    rng = np.random.default_rng(5)
    for h, w, gx, gy, slope in [(71, 113, 8, 8, 0.04), (48, 50, 5, 2, 0.01), (25, 31, 2, 6, 0.2)]:
        g = rng.integers(0, 256, (h, w), dtype=np.uint8)
        np.testing.assert_array_equal(L.adapthisteq(g, slope, gx, gy), L.ref_adapthisteq(lib, g, slope, gx, gy))
        e = rng.integers(0, 256, (h, w), dtype=np.uint8)
        np.testing.assert_array_equal(L.combine_results(g, e, 100), L.ref_combine_results(lib, g, e, 100))
        f0 = rng.integers(0, 256, (h, w, 3), dtype=np.uint8)
        f1 = np.where(rng.random((h, w, 1)) < 0.2, rng.integers(0, 256, (h, w, 3), dtype=np.uint8), f0)
        for jump in (1, 2, 5):
            a, b = L.speaker_detection(f1, f0, 20, jump), L.ref_speaker_detection(lib, f1, f0, 20, jump)
            assert a[0] == b[0]
            for x, y in zip(a[1:], b[1:]):
                np.testing.assert_array_equal(x, y)
