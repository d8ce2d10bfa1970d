"""CPU tests: the PNG container restated in oracle/png_oracle.py is a valid file for the reference's reader (cv2.imdecode,
helper.py:31) and for zlib, and am_png1_size (host-only helper of the C ABI) agrees with it."""
import zlib

import numpy as np
import pytest

from oracle import png_oracle as PO


@pytest.mark.parametrize("hw", [(1080, 1920), (37, 101), (5, 7), (300, 8), (1, 1), (273, 7285)])
def test_png1_roundtrip_through_the_reference_reader(hw):
    h, w = hw
    m = (np.random.default_rng(h + w).random((h, w)) < 0.4).astype(np.uint8) * 255
    png = PO.png1(m)
    assert len(png) == PO.size(w, h)
    np.testing.assert_array_equal(PO.decode(np.frombuffer(png, np.uint8)), m)
    idat = png[8 + 25 + 8:-12 - 4]
    assert len(zlib.decompress(idat)) == h * (1 + (w + 7) // 8)


def test_c_abi_size_helper():
    from lecturemath_b200 import _lib
    lib = _lib.load()
    for (w, h) in [(1920, 1080), (101, 37), (7, 5), (8, 300), (1, 1), (3840, 2160)]:
        assert lib.am_png1_size(w, h) == PO.size(w, h)
