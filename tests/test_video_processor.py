"""The frame source (lecturemath_b200/video_processor.py) against the unmodified reference's VideoProcessor.

CPU: the frames the worker receives -- which decoded frame, in which order, with which frame_time / rel_time / frame_idx /
last_frame / shape -- for seeded synthetic videos equal tests/golden/video_sampling.json, captured from the reference by
oracle/gen_golden_video.py (fps sampling, "all frames", the skipped first sample, the grab / seek probe, limit, several files,
forced resolution incl. a second file of another size and frame rate).  Videos are rewritten here with cv2.VideoWriter (MJPG);
frames carry their index as a block pattern, so the comparison does not depend on codec bytes.
GPU: with this package's batching binarizer as the worker the forced resize runs on the device."""
import json
import os

import numpy as np
import pytest

from oracle import gen_golden_video as GV
from oracle import resize_oracle as RO
from tests.conftest import GOLDEN


@pytest.fixture(scope="module")
def videos(tmp_path_factory):
    d = tmp_path_factory.mktemp("videos")
    try:
        return GV.write_videos(str(d))
    except RuntimeError as e:
        pytest.skip(str(e))


@pytest.mark.parametrize("i", range(len(GV.CASES)))
def test_sampling_equals_the_reference(videos, i):
    from lecturemath_b200.video_processor import VideoProcessor
    with open(os.path.join(GOLDEN, "video_sampling.json")) as f:
        ref = json.load(f)["cases"][i]
    got = GV.run_case(VideoProcessor, videos, GV.CASES[i])
    assert got["size"] == ref["size"] and got["finalized"] is True
    assert len(got["log"]) == len(ref["log"])
    for a, b in zip(got["log"], ref["log"]):
        assert a == b, (a, b)


def test_resolution_mismatch_without_forcing_raises(videos):
    from lecturemath_b200.video_processor import VideoProcessor
    vp = VideoProcessor([videos["a"], videos["c"]], 5)
    with pytest.raises(Exception, match="same resolution"):
        vp.doProcessing(GV.Recorder(), 0, False, True)


@pytest.mark.parametrize("shape", [((1080, 1920), (720, 1280)), ((2160, 3840), (1080, 1920)), ((144, 256), (72, 128)), ((300, 500), (300, 500))])
def test_linear_oracle_is_opencv_on_downscales(shape):
    """oracle/resize_oracle.cv2_linear_u8 (OpenCV's documented 8-bit INTER_LINEAR) is bit-identical to the installed cv2.resize
    whenever that build runs its own generic code: every down-scale (and the identity)."""
    cv2 = pytest.importorskip("cv2")
    (ih, iw), (oh, ow) = shape
    src = np.random.default_rng(ih + ow).integers(0, 256, (ih, iw, 3), dtype=np.uint8)
    np.testing.assert_array_equal(RO.cv2_linear_u8(src, ow, oh), cv2.resize(src, (ow, oh)))
    np.testing.assert_array_equal(RO.cv2_linear_u8(src[:, :, 0], ow, oh), cv2.resize(np.ascontiguousarray(src[:, :, 0]), (ow, oh)))


@pytest.mark.parametrize("shape", [((180, 320), (1080, 1920)), ((720, 1280), (1080, 1920)), ((90, 160), (135, 200)), ((144, 256), (270, 480))])
def test_linear_oracle_vs_opencv_on_upscales(shape):
    """Up-scales: this OpenCV build (4.13.0, IPP 2022.2) leaves its generic code there and the result differs from the documented
    fixed-point algorithm by ONE grey level on a few pixels per thousand (measured 0.05 % - 0.3 %; cv2.ipp.setUseIPP(False) does not
    change it).  Stated tolerance: |difference| <= 1 everywhere, on <= 0.5 % of the samples."""
    cv2 = pytest.importorskip("cv2")
    (ih, iw), (oh, ow) = shape
    src = np.random.default_rng(ih * 3 + ow).integers(0, 256, (ih, iw, 3), dtype=np.uint8)
    d = np.abs(RO.cv2_linear_u8(src, ow, oh).astype(int) - cv2.resize(src, (ow, oh)).astype(int))
    assert d.max() <= 1 and (d > 0).mean() <= 5e-3


@pytest.mark.gpu
@pytest.mark.parametrize("shape", [((1080, 1920), (720, 1280)), ((180, 320), (1080, 1920)), ((2160, 3840), (1080, 1920)), ((144, 256), (270, 480)),
                                   ((37, 101), (90, 55)), ((5, 7), (1, 1))])
def test_device_linear_resize_is_bit_exact_vs_oracle(shape):
    import ctypes
    import torch
    from lecturemath_b200 import _lib
    (ih, iw), (oh, ow) = shape
    rng = np.random.default_rng(ih * 7 + ow)
    for ch in (3, 1):
        src = rng.integers(0, 256, (2, ih, iw, ch), dtype=np.uint8)
        d_in = torch.from_numpy(src).cuda()
        d_out = torch.empty((2, oh, ow, ch), dtype=torch.uint8, device="cuda")
        _lib.check(_lib.lib().am_resize_linear_u8(d_in.data_ptr(), 2, ih, iw, ch, oh, ow, d_out.data_ptr(),
                                                 ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)), "am_resize_linear_u8")
        got = d_out.cpu().numpy()
        for f in range(2):
            ref = RO.cv2_linear_u8(src[f] if ch == 3 else src[f, :, :, 0], ow, oh)
            np.testing.assert_array_equal(got[f] if ch == 3 else got[f, :, :, 0], ref)


@pytest.mark.gpu
@pytest.mark.parametrize("batch", [1, 4])
def test_video_to_masks_with_forced_resolution_on_the_device(videos, golden, batch):
    """VideoProcessor -> FCN_LectureNet_Binarizer(batch) with a forced resolution: the worker receives the decoded 256x144 / 320x180
    frames, resizes them on the device and binarizes; labels (times, indices) equal the reference's log and every mask equals the one
    obtained from the oracle-resized frame through the same network (bit-exact: same resize arithmetic, same plan)."""
    import torch
    from lecturemath_b200.fcn_binarizer_worker import FCN_LectureNet_Binarizer
    from lecturemath_b200.helper import Helper
    from lecturemath_b200.video_processor import VideoProcessor
    from tests.test_fcn_host_logic import golden_net
    net = golden_net("tiny", golden("fcn_forward.npz")).cuda()
    with open(os.path.join(GOLDEN, "video_sampling.json")) as f:
        ref = json.load(f)["cases"][7]                                 # (["a", "c"], 5 fps, forced 320x180)
    vp = VideoProcessor([videos["a"], videos["c"]], 5)
    vp.force_resolution(320, 180)
    worker = FCN_LectureNet_Binarizer(net, batch=batch, keep_others=False)
    vp.doProcessing(worker, 0, False, True)
    assert worker.frame_indices == [e[3] for e in ref["log"]]
    np.testing.assert_allclose(worker.frame_times, [e[1] for e in ref["log"]], atol=1e-3)
    masks = Helper.decompress_binary_images(worker.compressed_frames)
    assert len(masks) == len(ref["log"]) and all(m.shape == (180, 320) for m in masks)
    # the same frames, resized by the oracle on the host, through a worker that gets them at the final size
    class Collect(GV.Recorder):
        def handleFrame(self, frame, last_frame, v_index, abs_time, rel_time, abs_frame_idx):
            self.log.append(frame.copy())
    raw = VideoProcessor([videos["a"], videos["c"]], 5)               # no forcing: frames as decoded (file c would raise -> one file at a time)
    frames = []
    for name in ("a", "c"):
        c = Collect()
        VideoProcessor([videos[name]], 5).doProcessing(c, 0, False, True)
        frames.append(c.log)
    # (per-file runs skip each file's first sample; the joint run only skips the very first one)
    first_c = None
    import cv2
    cap = cv2.VideoCapture(videos["c"])
    for _ in range(int(25.0 / 5)):
        ok, first_c = cap.read()
    cap.release()
    sized = [f for f in frames[0]] + [RO.cv2_linear_u8(f, 320, 180) for f in [first_c] + frames[1]]
    assert len(sized) == len(masks)
    direct = FCN_LectureNet_Binarizer(net, batch=batch, keep_others=False)
    direct.initialize(320, 180)
    for i, fr in enumerate(sized):
        direct.handleFrame(fr, None, 0, 0.0, 0.0, i)
    direct.finalize()
    for a, b in zip(masks, Helper.decompress_binary_images(direct.compressed_frames)):
        np.testing.assert_array_equal(np.asarray(a), np.asarray(b))
