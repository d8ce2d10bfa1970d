"""GPU parity tests for the > 2.5 MP branch (csrc/resize.cu through the C ABI): am_lanczos_resize_u8 == Pillow LANCZOS,
am_bits_resize_nearest == cv2 INTER_NEAREST (both bit-exact, against the oracle and the golden vectors), and the whole
large-frame path -- worker drop-in against the unmodified reference's output, fused pipeline against the oracle."""
import ctypes
import hashlib

import numpy as np
import pytest
import torch

from oracle import cc_oracle as CO, fcn_oracle as FO, resize_oracle as RO
from oracle.gen_golden_resize import LANCZOS_CASES, NEAREST_CASES
from tests.test_fcn_host_logic import golden_net

pytestmark = pytest.mark.gpu
MASK_TOL = 1e-3          # BASELINE.json north_star: mask disagreement <= 0.1 % of pixels (bf16 FCN vs fp32 reference)


def _lanczos_gpu(img, ow, oh, offset=0):
    """img: uint8 (B, H, W, C) -> (B, oh, ow, C) through am_lanczos_resize_u8; offset: misalign the input pointer."""
    from lecturemath_b200 import _lib
    lib = _lib.lib()
    b, h, w, c = img.shape
    raw = torch.empty(img.size + offset + 8, dtype=torch.uint8, device="cuda")
    raw[offset:offset + img.size] = torch.from_numpy(img.reshape(-1)).cuda()
    out = torch.zeros((b, oh, ow, c), dtype=torch.uint8, device="cuda")
    _lib.check(lib.am_lanczos_resize_u8(raw.data_ptr() + offset, b, h, w, c, oh, ow, out.data_ptr(),
                                        ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)), "am_lanczos_resize_u8")
    torch.cuda.synchronize()
    return out.cpu().numpy()


def _nearest_gpu(masks, ow, oh):
    """masks: uint8 (B, h, w) 0/255 -> (B, oh, ow) through pack -> am_bits_resize_nearest -> unpack."""
    from lecturemath_b200 import _lib
    from lecturemath_b200.cc_engine import CCEngine
    lib = _lib.lib()
    b, h, w = masks.shape
    src, dst = CCEngine(w, h, b), CCEngine(ow, oh, b)
    bits = src.pack(torch.from_numpy(masks).cuda())
    out = torch.full((b, oh, lib.am_words_per_row(ow)), -1, dtype=torch.int32, device="cuda")
    _lib.check(lib.am_bits_resize_nearest(bits.data_ptr(), b, h, w, oh, ow, out.data_ptr(),
                                          ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)), "am_bits_resize_nearest")
    torch.cuda.synchronize()
    words = out.cpu().numpy().view(np.uint32)
    assert not (words[:, :, (ow + 31) // 32:] != 0).any(), "padding words must stay zero"
    if ow % 32:
        assert not (words[:, :, ow // 32] >> (ow % 32)).any(), "padding bits must stay zero"
    return dst.unpack(out).cpu().numpy()


@pytest.mark.parametrize("i", range(len(LANCZOS_CASES)))
def test_lanczos_vs_pillow_golden(golden, i):
    z = golden("resize.npz")
    img, ref = z["lanczos_in_%d" % i], z["lanczos_out_%d" % i]
    for off in (0, 1, 3):
        got = _lanczos_gpu(img[None], ref.shape[1], ref.shape[0], off)
        np.testing.assert_array_equal(got[0], ref)


@pytest.mark.parametrize("hw", [(270, 480), (271, 481), (64, 1001), (333, 777), (200, 200)])
def test_lanczos_vs_oracle_batched(hw):
    h, w = hw
    rng = np.random.default_rng(h * 7 + w)
    img = rng.integers(0, 256, (3, h, w, 3), dtype=np.uint8)
    img[1, ::3] = 255; img[1, 1::3] = 0                      # saturating rows in one frame
    got = _lanczos_gpu(img, int(w / 2), int(h / 2))
    for f in range(3):
        np.testing.assert_array_equal(got[f], RO.lanczos_resize(img[f], int(w / 2), int(h / 2)))
    same = _lanczos_gpu(img, w, int(h / 2))                  # one axis unchanged: Pillow skips that pass
    np.testing.assert_array_equal(same[0], RO.lanczos_resize(img[0], w, int(h / 2)))


def test_lanczos_4k_frame_vs_oracle():
    """BASELINE configs[4] size: 3840x2160 BGR -> 1920x1080, every byte equal to the Pillow restatement."""
    from lecturemath_b200 import synth
    frame = next(iter(synth.whiteboard_frames(1, 2160, 3840, seed=5, chalk=True)))
    got = _lanczos_gpu(frame[None], 1920, 1080)[0]
    np.testing.assert_array_equal(got, RO.lanczos_resize(frame, 1920, 1080))
    const = np.full((1, 2160, 3840, 3), 201, dtype=np.uint8)                   # size-independent property: constants are preserved
    assert (_lanczos_gpu(const, 1920, 1080) == 201).all()


@pytest.mark.parametrize("i", range(len(NEAREST_CASES)))
def test_nearest_vs_opencv_golden(golden, i):
    z = golden("resize.npz")
    m, ref = z["nearest_in_%d" % i], z["nearest_out_%d" % i]
    np.testing.assert_array_equal(_nearest_gpu(m[None], ref.shape[1], ref.shape[0])[0], ref)


@pytest.mark.parametrize("shape", [(1080, 1920, 2160, 3840), (1080, 1920, 2161, 3841), (650, 1000, 1300, 2000), (650, 1000, 1301, 2001)])
def test_nearest_full_size_vs_oracle(shape):
    sh, sw, dh, dw = shape
    rng = np.random.default_rng(3)
    m = (rng.random((2, sh, sw)) < 0.3).astype(np.uint8) * 255
    got = _nearest_gpu(m, dw, dh)
    for f in range(2):
        np.testing.assert_array_equal(got[f], RO.nearest_resize(m[f], dw, dh))


def test_worker_on_2p6mp_frame_vs_reference_golden(golden):
    """FCN_LectureNet_Binarizer.handleFrame on a 2000x1300 frame: the device LANCZOS image equals the one the unmodified
    reference fed its FCN (sha256), the returned full-size ink mask differs from the reference's in <= 0.1 % of pixels."""
    from lecturemath_b200 import synth
    from lecturemath_b200.fcn_binarizer_worker import FCN_LectureNet_Binarizer
    z = golden("resize.npz")
    h, w = (int(v) for v in z["big_shape"])
    frame = next(iter(synth.whiteboard_frames(1, h, w, seed=int(z["big_seed"]))))
    if hashlib.sha256(frame.tobytes()).digest() != z["big_frame_sha256"].tobytes():
        pytest.skip("synthetic frame generator differs on this host; golden frame not reproducible")
    net = golden_net("tiny", golden("fcn_forward.npz")).cuda()
    worker = FCN_LectureNet_Binarizer(net)
    worker.initialize(w, h)
    worker.handleFrame(frame, None, 0, 0.0, 0.0, 0)
    plan = net.plan(1, 650, 1000)
    small_bgr = plan.frames[0].cpu().numpy()
    assert hashlib.sha256(np.ascontiguousarray(small_bgr[:, :, ::-1]).tobytes()).digest() == z["big_small_rgb_sha256"].tobytes()
    ref_ink = np.unpackbits(z["big_ink_bits"], axis=-1)[:, :w].astype(np.uint8) * 255
    assert worker.last_binary.shape == (h, w) and worker.last_text.shape == (h, w) and worker.last_rec.shape == (h, w, 3)
    assert (worker.last_binary != ref_ink).mean() <= MASK_TOL
    # the mask is a pure 2x nearest upscale of the FCN-size mask
    np.testing.assert_array_equal(worker.last_binary, RO.nearest_resize(worker.last_binary[::2, ::2], w, h))
    # binarize() drop-in on the same frame (PIL in, reference polarity out)
    from PIL import Image
    binary = net.binarize(Image.fromarray(np.ascontiguousarray(frame[:, :, ::-1])), force_binary=True)
    np.testing.assert_array_equal(255 - binary, worker.last_binary)


def test_fused_pipeline_on_large_frames_vs_oracle(golden):
    """ContentExtractor at 2000x1300 (FCN at 1000x650, CC stage at full size): masks vs the fp32 oracle path
    (PIL LANCZOS + torch fp32 + cv2 NEAREST), CC rows / tempo_count bit-exact vs the oracle on the GPU's masks."""
    from lecturemath_b200 import synth
    from lecturemath_b200.pipeline import ContentExtractor
    h, w, n = 1300, 2000, 3
    net = golden_net("tiny", golden("fcn_forward.npz")).cuda()
    frames = np.stack(list(synth.whiteboard_frames(n, h, w, seed=21)))
    ex = ContentExtractor(net, w, h, 0.85, 0.85, 85, batch=n, device="cuda:0")
    assert ex.large.active and (ex.plan.H, ex.plan.W) == (650, 1000)
    rows = ex.process_batch(frames)
    masks = ex.masks_host()
    assert masks.shape == (n, h, w)
    sd = net.state_dict()
    est = CO.StabilityOracle(w, h, 0.85, 0.85, 85)
    for f in range(n):
        ink_ref, _, _ = FO.handle_frame(sd, frames[f])
        assert (ink_ref != masks[f]).mean() <= MASK_TOL
        est.add_frame(masks[f])
        ref = np.array(est.frame_table(f), dtype=np.int64).reshape(-1, 7)
        np.testing.assert_array_equal(rows[f].astype(np.int64), ref)
    assert ex.est.state()["tempo_count"] == est.tempo_count


def test_batching_worker_on_large_frames_equals_fused_pipeline(golden):
    """The asynchronous worker on frames above 2.5 MP (device LANCZOS halving, FCN at half size, NEAREST mask upscale, per-slot mask
    copies for the PNG writer) with an estimator attached: the files decode to the masks the fused ContentExtractor produces for the
    same batches, and the estimator reaches the same rows / tempo_count."""
    from lecturemath_b200 import synth
    from lecturemath_b200.cc_stability_estimator import CCStabilityEstimator
    from lecturemath_b200.fcn_binarizer_worker import FCN_LectureNet_Binarizer
    from lecturemath_b200.helper import Helper
    from lecturemath_b200.pipeline import ContentExtractor
    h, w, b, n = 1300, 2000, 2, 5
    net = golden_net("tiny", golden("fcn_forward.npz")).cuda()
    frames = np.stack(list(synth.whiteboard_frames(n + 1, h, w, seed=23)))
    ex = ContentExtractor(net, w, h, 0.85, 0.85, 85, batch=b, device="cuda:0")
    ref_rows, ref_masks = [], []
    for s in range(0, n + 1, b):                                        # three full batches (the last frame only pads the third)
        ref_rows += ex.process_batch(frames[s:s + b])
        ref_masks.append(ex.masks_host())
    ref_masks = np.concatenate(ref_masks)
    est = CCStabilityEstimator(w, h, 0.85, 0.85, 85, max_batch=b)
    worker = FCN_LectureNet_Binarizer(net, batch=b, estimator=est)
    worker.initialize(w, h)
    for i in range(n):                                                 # five frames: the last batch is partial
        worker.handleFrame(frames[i], None, 0, 40.0 * i, 40.0 * i, i)
    worker.finalize()
    masks = Helper.decompress_binary_images(worker.compressed_frames)
    assert len(masks) == n and worker.frame_indices == list(range(n))
    for i in range(n):
        np.testing.assert_array_equal(np.asarray(masks[i]), ref_masks[i])
    got = [[(u, cc.cc_id + 1, int(cc.min_x), int(cc.max_x), int(cc.min_y), int(cc.max_y), int(cc.size)) for u, cc in fr] for fr in est.cc_idx_per_frame]
    assert got == [[tuple(int(v) for v in r) for r in ref_rows[i]] for i in range(n)]
