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


def _encode(masks):
    from lecturemath_b200.cc_engine import CCEngine
    from lecturemath_b200.wire import encode_png_frames
    n, h, w = masks.shape
    bits = CCEngine(w, h, n).pack(torch.from_numpy(masks).cuda())
    return encode_png_frames(bits, w, h)


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
    raw = zlib.decompress(idat)                                        # verifies Adler-32 and the stored-block framing
    assert len(raw) == h * (1 + (w + 7) // 8)


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
    for mode in ("device", "cv2"):
        worker = FCN_LectureNet_Binarizer(net, png=mode)
        worker.initialize(250, 180)
        for i, fr in enumerate(frames):
            worker.handleFrame(fr, None, 0, 33.3 * i, 33.3 * i, i)
        out[mode] = Helper.decompress_binary_images(worker.compressed_frames)
        assert all(isinstance(r, np.ndarray) and r.dtype == np.uint8 for r in worker.compressed_frames)
    for a, b in zip(out["device"], out["cv2"]):
        np.testing.assert_array_equal(a, b)
        assert a.shape == (180, 250) and set(np.unique(a)) <= {0, 255}
