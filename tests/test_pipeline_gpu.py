"""GPU tests of the fused pipelines (through the C ABI): ContentExtractor (synchronous) and StreamingExtractor (two
streams, asynchronous hand-off buffers) must produce identical, oracle-exact CC rows on the masks the FCN emits."""
import os

import numpy as np
import pytest
import torch

from oracle import cc_oracle as CO
from tests.conftest import GOLDEN

pytestmark = pytest.mark.gpu


def _tiny_net():
    from lecturemath_b200.configuration import Configuration
    from lecturemath_b200.fcn_lecturenet import FCN_LectureNet
    torch.manual_seed(0)
    return FCN_LectureNet.CreateFromConfig(Configuration.from_file(os.path.join(GOLDEN, "fcn_tiny.conf")), 3, False).eval()


def test_streaming_equals_synchronous_equals_oracle():
    from lecturemath_b200 import synth
    from lecturemath_b200.pipeline import ContentExtractor, StreamingExtractor
    h, w, b, steps = 180, 256, 4, 5
    frames = np.stack(list(synth.whiteboard_frames(b * steps, h, w, seed=3)))
    net = _tiny_net()
    ex = ContentExtractor(net, w, h, 0.85, 0.85, 85, batch=b, device="cuda:0")
    sync_rows, masks = [], []
    for s in range(steps):
        sync_rows += ex.process_batch(frames[s * b:(s + 1) * b])
        masks.append(ex.masks_host())
    masks = np.concatenate(masks)
    sx = StreamingExtractor(net, w, h, 0.85, 0.85, 85, batch=b, device="cuda:0")
    pinned = torch.from_numpy(frames).pin_memory()
    stream_rows = []
    for s in range(steps):
        sx.submit(pinned[s * b:(s + 1) * b], last=(s == steps - 1))
        if s >= 1:
            stream_rows += sx.collect(s - 1)
    stream_rows += sx.collect(steps - 1)
    state = sx.finish()
    est = CO.StabilityOracle(w, h, 0.85, 0.85, 85)
    for f in range(b * steps):
        est.add_frame(masks[f])
        ref = np.array(est.frame_table(f), dtype=np.int64).reshape(-1, 7)
        np.testing.assert_array_equal(sync_rows[f].astype(np.int64), ref)
        np.testing.assert_array_equal(stream_rows[f].astype(np.int64), ref)
    assert state["tempo_count"] == est.tempo_count and state["n_unique"] == len(est.unique_cc_objects)
    assert state["img_idx"] == b * steps
    # a second video through the same object starts from a clean temporal state
    sx.reset()
    sx.submit(pinned[:b], last=True)
    again = sx.collect(0)
    for f in range(b):
        np.testing.assert_array_equal(again[f], stream_rows[f])


def test_async_handoff_buffer_roundtrip(golden):
    """am_est_export_dev / am_est_import_dev (no host sync): split a run at frame k across two estimators."""
    from lecturemath_b200.cc_engine import CCEngine, Estimator
    from tests.conftest import unpack_masks
    z = golden("cc_stability.npz")
    masks = unpack_masks(z, "blobs_gap6")
    n, h, w = masks.shape
    r, p, gap = z["blobs_gap6_params"]
    ref = z["blobs_gap6_per_frame"]
    dev = torch.from_numpy(masks).cuda()
    for k, cap in ((1, 1 << 16), (23, 1 << 16), (40, 1 << 20)):
        eng = CCEngine(w, h, n)
        eng.label(eng.pack(dev), sync=False)
        a = Estimator(w, h, float(r), float(p), int(gap))
        a.add_frames(eng, 0, k)
        buf = torch.zeros(cap, dtype=torch.int32, device="cuda")
        a.export_dev(buf)
        b = Estimator(w, h, float(r), float(p), int(gap))
        b.import_dev(buf)
        b.add_frames(eng, k, n - k)
        eng.read_counts()
        rows, offs = eng.packed_rows(n)
        rows, offs = rows.cpu().numpy(), offs.cpu().numpy()
        got = np.array([(t,) + tuple(int(v) for v in row[:7]) for t in range(n) for row in rows[offs[t]:offs[t + 1]]], dtype=np.int64).reshape(-1, 8)
        np.testing.assert_array_equal(got, ref)
        assert b.state()["tempo_count"] == int(z["blobs_gap6_tempo"])
    # a hand-off buffer that is too small must fail loudly, not silently drop uniques
    small = torch.zeros(64, dtype=torch.int32, device="cuda")
    a.export_dev(small)
    c = Estimator(w, h, float(r), float(p), int(gap))
    c.import_dev(small)
    from lecturemath_b200._lib import AccessMathB200Error
    with pytest.raises(AccessMathB200Error):
        c.state()
