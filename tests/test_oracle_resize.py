"""CPU tests: the resize oracle (Pillow LANCZOS / OpenCV INTER_NEAREST restated in numpy, oracle/resize_oracle.py) against
the golden vectors captured from Pillow, OpenCV and the unmodified reference (tests/golden/resize.npz), and directly against
the two libraries where they are importable.  Bit-exact."""
import hashlib

import numpy as np
import pytest

from oracle import resize_oracle as RO
from oracle.gen_golden_resize import LANCZOS_CASES, NEAREST_CASES


@pytest.mark.parametrize("i", range(len(LANCZOS_CASES)))
def test_lanczos_matches_pillow_golden(golden, i):
    z = golden("resize.npz")
    img, ref = z["lanczos_in_%d" % i], z["lanczos_out_%d" % i]
    got = RO.lanczos_resize(img, ref.shape[1], ref.shape[0])
    np.testing.assert_array_equal(got, ref)


@pytest.mark.parametrize("i", range(len(NEAREST_CASES)))
def test_nearest_matches_opencv_golden(golden, i):
    z = golden("resize.npz")
    m, ref = z["nearest_in_%d" % i], z["nearest_out_%d" % i]
    np.testing.assert_array_equal(RO.nearest_resize(m, ref.shape[1], ref.shape[0]), ref)


def test_reference_halved_image_of_the_2p6mp_frame(golden):
    """The image the unmodified reference fed its FCN for the seeded 2000x1300 frame == oracle LANCZOS of that frame."""
    from lecturemath_b200 import synth
    z = golden("resize.npz")
    h, w = (int(v) for v in z["big_shape"])
    frame = next(iter(synth.whiteboard_frames(1, h, w, seed=int(z["big_seed"]))))
    if hashlib.sha256(frame.tobytes()).digest() != z["big_frame_sha256"].tobytes():
        pytest.skip("synthetic frame generator differs on this host (OpenCV anti-aliasing); golden frame not reproducible")
    small = RO.halve_until_fits(np.ascontiguousarray(frame[:, :, ::-1]))
    assert small.shape == (650, 1000, 3)
    np.testing.assert_array_equal(small[[0, 1, 324, 648, 649]], z["big_small_rgb_rows"])
    assert hashlib.sha256(small.tobytes()).digest() == z["big_small_rgb_sha256"].tobytes()


def test_against_installed_pillow_and_opencv():
    PIL_Image = pytest.importorskip("PIL.Image")
    cv2 = pytest.importorskip("cv2")
    rng = np.random.default_rng(5)
    for (h, w) in [(270, 480), (271, 481), (64, 1001), (540, 960)]:
        img = rng.integers(0, 256, (h, w, 3), dtype=np.uint8)
        ref = np.asarray(PIL_Image.fromarray(img).resize((int(w / 2), int(h / 2)), PIL_Image.LANCZOS))
        np.testing.assert_array_equal(RO.lanczos_resize(img, int(w / 2), int(h / 2)), ref)
        m = (rng.random((int(h / 2), int(w / 2))) < 0.5).astype(np.uint8) * 255
        np.testing.assert_array_equal(RO.nearest_resize(m, w, h), cv2.resize(m, (w, h), interpolation=cv2.INTER_NEAREST))


def test_guard_sizes():
    assert RO.fcn_size(1920, 1080) == [(1920, 1080)]
    assert RO.fcn_size(3840, 2160) == [(3840, 2160), (1920, 1080)]
    assert RO.fcn_size(7680, 4320) == [(7680, 4320), (3840, 2160), (1920, 1080)]
    assert RO.fcn_size(2001, 1301) == [(2001, 1301), (1000, 650)]
    from lecturemath_b200.large_frames import working_sizes
    for wh in [(1920, 1080), (3840, 2160), (7680, 4320), (2001, 1301), (2704, 1520)]:
        assert working_sizes(*wh) == RO.fcn_size(*wh)


def test_c_abi_working_size_helper():
    """am_fcn_working_size is a host-only helper of the C ABI (no device needed)."""
    import ctypes
    from lecturemath_b200 import _lib
    lib = _lib.load()
    for wh in [(1920, 1080), (3840, 2160), (7680, 4320), (2001, 1301)]:
        ow, oh = ctypes.c_int(0), ctypes.c_int(0)
        n = lib.am_fcn_working_size(wh[0], wh[1], ctypes.byref(ow), ctypes.byref(oh))
        sizes = RO.fcn_size(*wh)
        assert n == len(sizes) - 1 and (ow.value, oh.value) == sizes[-1]
