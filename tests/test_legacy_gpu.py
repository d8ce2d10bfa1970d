"""GPU parity tests for the four legacy accessmath_lib exports (csrc/legacy_ops.cu), called through the C ABI with the
reference's own ctypes calling convention (host NumPy buffers), against the golden vectors captured from the
reference's compiled C and the NumPy oracle.  Bar: bit-exact, fp64 outputs included."""
import ctypes

import numpy as np
import pytest

from oracle import legacy_oracle as L
from oracle.gen_golden_legacy import legacy_inputs

pytestmark = pytest.mark.gpu
INPUTS = legacy_inputs()


def _lib():
    from lecturemath_b200 import _lib
    return _lib.lib()


def _p(a):
    return a.ctypes.data_as(ctypes.c_void_p)


def run_cuda(lib, name, d):
    # L.ref_* are thin ctypes callers written for the reference's .so; our library exports the same symbols
    if name.startswith("ahe"):
        return L.ref_adapthisteq(lib, d["gray"], d["slope"], d["gx"], d["gy"])
    if name.startswith("comb"):
        return L.ref_combine_results(lib, d["board"], d["eq"], d["thr"])
    t, b, a, dv = L.ref_speaker_detection(lib, d["frame"], d["last"], d["thr"], d["jump"])
    return np.concatenate([b, a, dv, [float(t)]])


@pytest.mark.parametrize("name", sorted(INPUTS))
def test_legacy_exports_vs_reference_golden(golden, name):
    lib = ctypes.CDLL(__import__("lecturemath_b200._lib", fromlist=["LIB_PATH"]).LIB_PATH)     # raw CDLL, as labeler.py:24 does
    _lib()
    z = golden("legacy_ops.npz")
    d = INPUTS[name]
    got = run_cuda(lib, name, d)
    np.testing.assert_array_equal(got, z[name])
    if name.startswith("ahe"):
        h, w = d["gray"].shape
        cdf = L.ref_region_cdf(lib, d["gray"], w // 5, w - 3, h // 4, h - 2, d["slope"])
        np.testing.assert_array_equal(cdf, z[name + "_cdf"])


def test_legacy_exports_full_size_vs_oracle():
    """1080p: adapthisteq with the reference's call parameters (binarizer.py:153), combine_results, speaker detection."""
    lib = ctypes.CDLL(__import__("lecturemath_b200._lib", fromlist=["LIB_PATH"]).LIB_PATH)
    _lib()
    rng = np.random.default_rng(11)
    h, w = 1080, 1920
    yy, xx = np.mgrid[0:h, 0:w]
    gray = np.clip(220 - 50.0 * xx / w - 30.0 * yy / h + rng.normal(0, 5, (h, w)), 0, 255).astype(np.uint8)
    gray[rng.random((h, w)) < 0.05] = 45
    eq = L.ref_adapthisteq(lib, gray, 0.04, 8, 8)
    np.testing.assert_array_equal(eq, L.adapthisteq(gray, 0.04, 8, 8))
    board = rng.integers(0, 256, (h, w), dtype=np.uint8)
    np.testing.assert_array_equal(L.ref_combine_results(lib, board, eq, 110), L.combine_results(board, eq, 110))
    f0 = rng.integers(0, 256, (h, w, 3), dtype=np.uint8)
    f1 = f0.copy(); f1[300:700, 500:900] = rng.integers(0, 256, (400, 400, 3), dtype=np.uint8)
    for jump in (1, 4):
        a, b = L.ref_speaker_detection(lib, f1, f0, 30, jump), L.speaker_detection(f1, f0, 30, jump)
        assert a[0] == b[0] and a[0] > 0
        for x, y in zip(a[1:], b[1:]):
            np.testing.assert_array_equal(x, y)


def test_device_pointer_variants_match_host_exports():
    import torch
    lib = _lib()
    rng = np.random.default_rng(2)
    h, w = 123, 257
    gray = rng.integers(0, 256, (h, w), dtype=np.uint8)
    d_gray = torch.from_numpy(gray).cuda(); d_out = torch.empty_like(d_gray)
    st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
    assert lib.am_adapthisteq_dev(d_gray.data_ptr(), w, h, 0.04, 4, 3, d_out.data_ptr(), st) == 0
    np.testing.assert_array_equal(d_out.cpu().numpy(), L.adapthisteq(gray, 0.04, 4, 3))
    d_cdf = torch.empty(256, dtype=torch.float64, device="cuda")
    assert lib.am_region_cdf_dev(d_gray.data_ptr(), w, h, 5, 200, 7, 100, 0.04, d_cdf.data_ptr(), st) == 0
    np.testing.assert_array_equal(d_cdf.cpu().numpy(), L.region_cdf(gray, 5, 200, 7, 100, 0.04))
    eq = rng.integers(0, 256, (h, w), dtype=np.uint8); d_eq = torch.from_numpy(eq).cuda()
    assert lib.am_combine_results_dev(d_gray.data_ptr(), d_eq.data_ptr(), w, h, 77, d_out.data_ptr(), st) == 0
    np.testing.assert_array_equal(d_out.cpu().numpy(), L.combine_results(gray, eq, 77))
    # unaligned planes take the scalar path
    flat_b = torch.from_numpy(np.concatenate([[0], gray.ravel()]).astype(np.uint8)).cuda()
    flat_e = torch.from_numpy(np.concatenate([[0], eq.ravel()]).astype(np.uint8)).cuda()
    flat_o = torch.empty_like(flat_b)
    assert lib.am_combine_results_dev(flat_b.data_ptr() + 1, flat_e.data_ptr() + 1, w, h, 77, flat_o.data_ptr() + 1, st) == 0
    np.testing.assert_array_equal(flat_o.cpu().numpy()[1:].reshape(h, w), L.combine_results(gray, eq, 77))
    f0 = rng.integers(0, 256, (h, w, 3), dtype=np.uint8)
    f1 = np.where(rng.random((h, w, 1)) < 0.1, 255 - f0, f0).astype(np.uint8)
    res = torch.empty(9, dtype=torch.float64, device="cuda")
    assert lib.am_speaker_detection_dev(torch.from_numpy(f1).cuda().data_ptr(), torch.from_numpy(f0).cuda().data_ptr(), w, h, 3, 40, 2,
                                        res.data_ptr(), st) == 0
    t, b, a, dv = L.speaker_detection(f1, f0, 40, 2)
    np.testing.assert_array_equal(res.cpu().numpy(), np.concatenate([b, a, dv, [float(t)]]))


def test_compute_binary_sums_equals_numpy_on_every_kind_of_frame():
    """VideoSegmenter.compute_binary_sums (R/AccessMath/preprocessing/content/video_segmenter.py:21-28: `binary.sum() / 255` per
    frame): lazy bit-packed frames (PNG scanlines and device words), decoded uint8 frames incl. the 254 values stage 03 can produce,
    odd sizes -- every entry bit-identical to the reference's expression, in the input order."""
    from lecturemath_b200.helper import Helper
    from lecturemath_b200.packed_mask import PackedMask
    from lecturemath_b200.video_segmenter import VideoSegmenter
    from oracle import png_oracle as PO
    rng = np.random.default_rng(12)
    frames, dense = [], []
    for k, (h, w) in enumerate([(180, 256), (180, 256), (37, 101), (180, 256), (1080, 1920), (37, 101), (180, 256)]):
        m = (rng.random((h, w)) < 0.1 * (k + 1)).astype(np.uint8) * 255
        if k == 3:
            m[5:9, 7:30] = 254                                         # a clean frame where two groups overlapped (uint8 wrap)
            frames.append(m)
        elif k % 3 == 0:
            frames.append(Helper.decompress_binary_images([np.frombuffer(PO.png1_deflate(m), np.uint8)])[0])    # scanline form
        elif k % 3 == 1:
            frames.append(PackedMask.from_dense(m))                    # word form
        else:
            frames.append(m)
        dense.append(m)
    frames.append(np.ones((10, 10), dtype=np.int32) * 255)             # not a decoded PNG: numpy's own reduction
    dense.append(frames[-1])
    got = VideoSegmenter.compute_binary_sums(frames, batch=2)
    ref = [b.sum() / 255 for b in dense]
    assert len(got) == len(ref)
    for a, b in zip(got, ref):
        assert float(a) == float(b) and isinstance(float(a), float)
