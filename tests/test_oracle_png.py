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


@pytest.mark.parametrize("hw,density", [((1080, 1920), 0.001), ((37, 101), 0.4), ((5, 7), 0.5), ((300, 8), 0.1), ((1, 1), 1.0),
                                        ((64, 4096), 0.02), ((273, 7285), 0.0), ((200, 300), 1.0), ((130, 1000), 0.5)])
def test_png1_deflate_is_a_valid_png_for_the_reference_reader(hw, density):
    """The compressed container (fixed-Huffman deflate, run-length matches, one block per 8 KB + empty stored blocks for byte
    alignment): accepted by zlib, decoded by cv2.imdecode to the mask, never larger than am_png1_capacity."""
    h, w = hw
    m = (np.random.default_rng(h * 7 + w).random((h, w)) < density).astype(np.uint8) * 255
    png = PO.png1_deflate(m)
    assert len(png) <= PO.capacity(w, h)
    np.testing.assert_array_equal(PO.decode(np.frombuffer(png, np.uint8)), m)
    idat_len = int.from_bytes(png[33:37], "big")
    raw = zlib.decompress(png[41:41 + idat_len])
    assert len(raw) == h * (1 + (w + 7) // 8)
    if density <= 0.001 and h * w > 100000:
        assert len(png) * 8 < PO.size(w, h)                              # sparse masks: at least 8x below the stored form


def test_capacity_helper_and_lazy_decode():
    """am_png1_capacity (host-only C helper) agrees with the restatement; Helper.decompress_binary_images returns the lazy bit-packed
    view for this package's files (stored and compressed) and plain arrays for anybody else's PNG -- same pixels either way."""
    import cv2
    from lecturemath_b200 import _lib
    from lecturemath_b200.helper import Helper
    from lecturemath_b200.packed_mask import PackedMask
    lib = _lib.load()
    for (w, h) in [(1920, 1080), (101, 37), (7, 5), (8, 300), (1, 1), (3840, 2160)]:
        assert lib.am_png1_capacity(w, h) == PO.capacity(w, h)
    rng = np.random.default_rng(5)
    for (h, w) in [(90, 133), (64, 64), (7, 9)]:
        m = (rng.random((h, w)) < 0.2).astype(np.uint8) * 255
        files = [np.frombuffer(PO.png1(m), np.uint8), np.frombuffer(PO.png1_deflate(m), np.uint8), cv2.imencode(".png", m)[1]]
        out = Helper.decompress_binary_images(files)
        assert isinstance(out[0], PackedMask) and isinstance(out[1], PackedMask) and isinstance(out[2], np.ndarray)
        for o in out:
            assert o.shape == (h, w)
            np.testing.assert_array_equal(np.asarray(o), m)
        np.testing.assert_array_equal(out[1][3:5, 2:9], m[3:5, 2:9])
        assert out[0].count_nonzero() == int((m != 0).sum())
        np.testing.assert_array_equal(out[1].words, PackedMask.from_dense(m).words)
        for o in Helper.decompress_binary_images(files, lazy=False):
            assert isinstance(o, np.ndarray)
