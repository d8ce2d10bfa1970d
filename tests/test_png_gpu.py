"""GPU tests for the device PNG writer (csrc/png.cu): byte-identical to the CPU restatement of the same
container (oracle/png_oracle.py), checksums accepted by zlib, and -- the property that matters for the drop-in -- decoded by the
reference's reader (cv2.imdecode, helper.py:31) to exactly the mask that went in."""
import struct
import zlib

import numpy as np
import pytest
import torch

from oracle import png_oracle as PO

pytestmark = pytest.mark.gpu


def _encode(masks, compress=False):
    from lecturemath_b200.cc_engine import CCEngine
    from lecturemath_b200.wire import encode_png_frames
    n, h, w = masks.shape
    bits = CCEngine(w, h, n).pack(torch.from_numpy(masks).cuda())
    return encode_png_frames(bits, w, h, compress=compress)


def _check_container(png, w, h):
    """Walk the chunks with zlib as the checksum authority."""
    assert png[:8] == PO.SIGNATURE
    pos, kinds, idat = 8, [], b""
    while pos < len(png):
        n, kind = struct.unpack(">I", png[pos:pos + 4])[0], png[pos + 4:pos + 8]
        data = png[pos + 8:pos + 8 + n]
        assert struct.unpack(">I", png[pos + 8 + n:pos + 12 + n])[0] == zlib.crc32(kind + data), kind
        kinds.append(kind)
        if kind == b"IDAT":
            idat += data
        pos += 12 + n
    assert kinds == [b"IHDR", b"IDAT", b"IEND"] and pos == len(png)
    raw = zlib.decompress(idat)                                        # verifies Adler-32 and the block framing
    assert len(raw) == h * (1 + (w + 7) // 8)


@pytest.mark.parametrize("hw", [(1080, 1920), (720, 1280), (2160, 3840), (37, 101), (5, 7), (300, 8), (273, 65535 // 9 + 3), (1, 1), (64, 4096)])
def test_deflate_png_bytes_checksums_and_reference_reader(hw):
    """The compressed writer (k_png1_deflate): byte-identical to its CPU restatement, valid for zlib, decoded by the reference's
    reader to the mask; noise, empty, full and sparse-stroke frames (the last one is what whiteboard video looks like)."""
    h, w = hw
    rng = np.random.default_rng(h * 31 + w)
    masks = (rng.random((4, h, w)) < 0.3).astype(np.uint8) * 255
    masks[1] = 0
    masks[2] = 255
    masks[3] = (rng.random((h, w)) < 0.002).astype(np.uint8) * 255
    files = _encode(masks, compress=True)
    for f in range(4):
        png = files[f].tobytes()
        _check_container(png, w, h)
        np.testing.assert_array_equal(PO.decode(files[f]), masks[f])   # cv2.imdecode(raw, IMREAD_GRAYSCALE)
        if h * w <= 1280 * 720 or f == 1:                              # (the Python restatement is slow on large noisy frames)
            assert png == PO.png1_deflate(masks[f]), "frame %d differs from the CPU restatement of the container" % f
    if h * w >= 100000:
        assert len(files[1]) * 50 < PO.size(w, h) and len(files[3]) * 4 < PO.size(w, h)


@pytest.mark.parametrize("hw", [(1080, 1920), (37, 101), (5, 7), (1, 1), (64, 4096)])
def test_deflate_png8_bytes_and_reference_reader(hw):
    """The 8-bit form of the writer (stage 03's clean frames where overlapping groups wrapped to 254): byte-identical to the CPU
    restatement, valid container, decoded by cv2.imdecode to the frame."""
    from lecturemath_b200.wire import PngEncoder
    h, w = hw
    rng = np.random.default_rng(h * 13 + w)
    frames = (rng.random((3, h, w)) < 0.02).astype(np.uint8) * 255
    frames[1][rng.random((h, w)) < 0.01] = 254
    frames[2] = rng.integers(0, 256, (h, w), dtype=np.uint8)
    files = PngEncoder(w, h, 3, compress=True, depth=8).encode(torch.from_numpy(frames).cuda())
    for f in range(3):
        png = files[f].tobytes()
        assert PO.png8_deflate(frames[f]) == png
        assert len(png) <= PO.capacity(w, h, 8)
        np.testing.assert_array_equal(PO.decode(files[f]), frames[f])


def test_scanlines_to_bits_roundtrip_on_device():
    """Decode half: file -> (host: zlib inflate, lazy PackedMask) -> am_png1_scanlines_to_bits -> the words the encoder started from."""
    from lecturemath_b200 import _lib
    from lecturemath_b200.cc_engine import CCEngine
    from lecturemath_b200.helper import Helper
    import ctypes
    rng = np.random.default_rng(9)
    for (h, w) in [(180, 250), (37, 101), (64, 4096), (5, 7)]:
        masks = (rng.random((3, h, w)) < 0.25).astype(np.uint8) * 255
        eng = CCEngine(w, h, 3)
        bits = eng.pack(torch.from_numpy(masks).cuda())
        from lecturemath_b200.wire import encode_png_frames
        lazy = Helper.decompress_binary_images(encode_png_frames(bits, w, h))
        scan = torch.from_numpy(np.stack([m.scan for m in lazy])).cuda()
        back = torch.zeros_like(bits)
        _lib.check(_lib.lib().am_png1_scanlines_to_bits(scan.data_ptr(), 3, h, w, back.data_ptr(), ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)),
                   "am_png1_scanlines_to_bits")
        assert torch.equal(back, bits)
        np.testing.assert_array_equal(np.stack([m.words for m in lazy]).view(np.int32), bits.cpu().numpy())


@pytest.mark.parametrize("hw", [(1080, 1920), (720, 1280), (2160, 3840), (37, 101), (5, 7), (300, 8), (273, 65535 // 9 + 3), (1, 1)])
def test_png_bytes_checksums_and_reference_reader(hw):
    h, w = hw
    rng = np.random.default_rng(h * 31 + w)
    masks = (rng.random((3, h, w)) < 0.3).astype(np.uint8) * 255
    masks[1] = 0
    masks[2] = 255
    files = _encode(masks)
    for f in range(3):
        png = files[f].tobytes()
        assert files[f].dtype == np.uint8 and files[f].ndim == 1 and len(png) == PO.size(w, h)
        _check_container(png, w, h)
        assert png == PO.png1(masks[f]), "frame %d differs from the CPU restatement of the container" % f
        np.testing.assert_array_equal(PO.decode(files[f]), masks[f])   # cv2.imdecode(raw, IMREAD_GRAYSCALE)


def test_worker_wire_format_device_equals_cv2(golden):
    """FCN_LectureNet_Binarizer: compressed_frames written on the device decode (through the drop-in Helper, i.e. the reference's
    reader) to the same masks as the cv2-encoded ones."""
    from lecturemath_b200 import synth
    from lecturemath_b200.fcn_binarizer_worker import FCN_LectureNet_Binarizer
    from lecturemath_b200.helper import Helper
    from tests.test_fcn_host_logic import golden_net
    net = golden_net("tiny", golden("fcn_forward.npz")).cuda()
    frames = list(synth.whiteboard_frames(3, 180, 250, seed=4))
    out = {}
    for mode in ("device", "device-stored", "cv2"):
        worker = FCN_LectureNet_Binarizer(net, png=mode)
        worker.initialize(250, 180)
        for i, fr in enumerate(frames):
            worker.handleFrame(fr, None, 0, 33.3 * i, 33.3 * i, i)
        out[mode] = Helper.decompress_binary_images(worker.compressed_frames)
        assert all(isinstance(r, np.ndarray) and r.dtype == np.uint8 for r in worker.compressed_frames)
    for a, b, c in zip(out["device"], out["cv2"], out["device-stored"]):
        np.testing.assert_array_equal(np.asarray(a), b)
        np.testing.assert_array_equal(np.asarray(c), b)
        assert a.shape == (180, 250) and set(np.unique(np.asarray(a))) <= {0, 255}


@pytest.mark.parametrize("n_frames,batch", [(11, 4), (8, 4), (3, 8)])
def test_batching_worker_and_lazy_estimator_equal_the_per_frame_protocol(golden, n_frames, batch):
    """The reference's calls, made asynchronous: handleFrame x n -> finalize -> decompress_binary_images -> add_frame x n ->
    finish_processing through the batching worker / the staging estimator: entries, indices and times in the per-frame order,
    partial last batches included; masks equal to the per-frame worker's up to near-threshold pixels (the network's tiling depends
    on the batch size: <= 0.1 %, the FCN bar); the staging estimator's state IDENTICAL to the frame-by-frame one on the same masks;
    and the worker that feeds an estimator directly (no PNG round trip) reaches that same state."""
    from lecturemath_b200 import synth
    from lecturemath_b200.cc_stability_estimator import CCStabilityEstimator
    from lecturemath_b200.fcn_binarizer_worker import FCN_LectureNet_Binarizer
    from lecturemath_b200.helper import Helper
    from tests.test_fcn_host_logic import golden_net
    net = golden_net("tiny", golden("fcn_forward.npz")).cuda()
    h, w = 180, 256
    frames = list(synth.whiteboard_frames(n_frames, h, w, seed=6))

    def run(worker):
        worker.initialize(w, h)
        for i, fr in enumerate(frames):
            worker.handleFrame(fr, None, 0, 33.3 * i, 33.3 * i, 100 + i)
        worker.finalize()
        return worker

    def state(est):
        return ([[(u, cc.cc_id, int(cc.min_x), int(cc.max_x), int(cc.min_y), int(cc.max_y), int(cc.size)) for u, cc in fr] for fr in est.cc_idx_per_frame],
                est.unique_cc_frames, est.tempo_count, est.img_idx, est.get_raw_cc_count())

    one = run(FCN_LectureNet_Binarizer(net, keep_others=False))
    fused_est = CCStabilityEstimator(w, h, 0.85, 0.85, 85, max_batch=batch)
    many = run(FCN_LectureNet_Binarizer(net, batch=batch, estimator=fused_est))
    assert many.frame_indices == one.frame_indices == [100 + i for i in range(n_frames)] and many.frame_times == one.frame_times
    assert len(many.compressed_frames) == n_frames
    masks_one = Helper.decompress_binary_images(one.compressed_frames, lazy=False)
    masks_many = Helper.decompress_binary_images(many.compressed_frames)
    dense_many = Helper.decompress_binary_images(many.compressed_frames, lazy=False)
    for a, b, c in zip(masks_one, masks_many, dense_many):
        np.testing.assert_array_equal(c, np.asarray(b))
        assert (a != c).mean() <= 1e-3
    np.testing.assert_array_equal(np.asarray(many.last_binary), dense_many[-1])
    ref = CCStabilityEstimator(w, h, 0.85, 0.85, 85, max_batch=1)
    for m in dense_many:
        ref.add_frame(m, True)                                           # dense arrays, one frame per launch sequence
    ref.finish_processing()
    lazy = CCStabilityEstimator(w, h, 0.85, 0.85, 85, max_batch=batch)
    for m in masks_many:
        lazy.add_frame(m, True)                                          # bit-packed lazy views, staged and batched
    lazy.finish_processing()
    assert state(lazy) == state(ref) == state(fused_est)
    for a, b in zip(lazy.unique_cc_objects, ref.unique_cc_objects):
        np.testing.assert_array_equal(a.img, b.img)
