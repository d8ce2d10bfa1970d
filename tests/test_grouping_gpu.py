"""GPU parity tests for stage 03 (csrc/grouping.cu + lecturemath_b200/cc_grouping.py): the CUDA drop-in
driven exactly like R/pre_ST3D_v3.0_03_cc_grouping.py:41-101 drives the reference estimator; every list, table and image
bit-exact against the outputs of the unmodified reference (tests/golden/cc_grouping.npz) and against the oracle."""
import pickle

import numpy as np
import pytest

from oracle import cc_oracle as CO
from oracle.gen_golden_grouping import RUNS, run_stage03
from oracle.grouping_oracle import GroupingOracle
from tests.conftest import unpack_masks
from tests.test_oracle_grouping import EXACT_KEYS, compare_with_golden

pytestmark = pytest.mark.gpu


def _stage02(zs, name):
    from lecturemath_b200.cc_stability_estimator import CCStabilityEstimator
    masks = unpack_masks(zs, name)
    r, p, gap = zs[name + "_params"]
    est = CCStabilityEstimator(masks.shape[2], masks.shape[1], float(r), float(p), int(gap))
    est.add_frames(masks)
    return est, masks


@pytest.mark.parametrize("name", sorted(RUNS))
@pytest.mark.parametrize("source", ["live", "unpickled"])
def test_stage03_vs_reference_golden(golden, name, source):
    """live: stage 03 reads the unique-CC tables stage 02 left in HBM (zero copy); unpickled: the estimator went through the
    tempo_stability pickle first, so the packed crops are re-uploaded."""
    zs, zg = golden("cc_stability.npz"), golden("cc_grouping.npz")
    est, _ = _stage02(zs, name)
    if source == "unpickled":
        est = pickle.loads(pickle.dumps(est, protocol=pickle.HIGHEST_PROTOCOL))
        assert not hasattr(est, "_est")
    split_gap, min_times, t_window, g_recall, img_t = zg[name + "/params"]
    res = run_stage03(est, int(split_gap), int(min_times), int(t_window), float(g_recall), float(img_t))
    compare_with_golden(zg, name, res)


def test_stage03_dense_1080p_vs_oracle():
    """BASELINE configs[3]-style masks at full size (dense glyphs, occluder): CUDA drop-in == oracle, every result."""
    from lecturemath_b200 import synth
    from lecturemath_b200.cc_stability_estimator import CCStabilityEstimator
    h, w, n = 1080, 1920, 10
    masks = np.stack(list(synth.glyph_masks(n, h, w, seed=2, churn=0.04)))
    est = CCStabilityEstimator(w, h, 0.85, 0.85, 85)
    est.add_frames(masks)
    stab = CO.StabilityOracle(w, h, 0.85, 0.85, 85)
    for m in masks:
        stab.add_frame(m)
    got = run_stage03(est, 3, 3, 5, 0.5, 0.5)
    ref = run_stage03(GroupingOracle(stab), 3, 3, 5, 0.5, 0.5)
    assert len(ref["stable"]) > 3000
    for k in EXACT_KEYS:
        assert ref[k].shape == got[k].shape and np.array_equal(ref[k], got[k]), k


def test_paint_wraparound_and_unaligned_rows():
    """am_paint_frames: three images over the same pixels give 253 (uint8 wrap-around of += 255), odd widths (rows not
    4-byte aligned) and the last pixels of the last frame."""
    from lecturemath_b200.cc_stability_estimator import CCStabilityEstimator
    from lecturemath_b200.connected_component import ConnectedComponent
    h, w = 37, 101
    est = CCStabilityEstimator(w, h, 0.85, 0.85, 85)
    rng = np.random.default_rng(0)
    ccs, ref = [], np.zeros((2, h, w), dtype=np.uint8)
    frames = []
    for k in range(40):
        x0, y0 = int(rng.integers(0, w - 1)), int(rng.integers(0, h - 1))
        x1, y1 = int(rng.integers(x0, min(w, x0 + 70))), int(rng.integers(y0, min(h, y0 + 30)))
        if k == 0:
            x0, y0, x1, y1 = 0, 0, w - 1, h - 1
        img = (rng.random((y1 - y0 + 1, x1 - x0 + 1)) < 0.7).astype(np.uint8) * 255
        ccs.append(ConnectedComponent(k, x0, x1, y0, y1, int((img > 0).sum()), img=img))
        f = k % 2
        frames.append(f)
        ref[f, y0:y1 + 1, x0:x1 + 1] += img
    got = est._paint(2, frames, ccs)
    np.testing.assert_array_equal(np.stack(got), ref)
    assert ref.max() == 255 and len(np.unique(ref)) > 4


@pytest.mark.parametrize("name", sorted(RUNS))
def test_keyframes_for_intervals_vs_reference_golden(golden, name):
    """KeyframeExtractor.GenerateFromST3DForIntervals with the overlap tests and the painting on the device: key-frames and sorted
    (time, box) lists equal the unmodified reference's (tests/golden/keyframes.npz) on four videos x five intervals."""
    from types import SimpleNamespace
    from lecturemath_b200.keyframe_extractor import KeyframeExtractor
    from oracle.gen_golden_keyframes import unpack_st3d
    from tests.test_oracle_keyframes import check, expected
    z = golden("keyframes.npz")
    n, h, w, segs, content, ref_times = expected(z, name)
    ages, images, bounds = unpack_st3d(z, name + "/")
    st3d = SimpleNamespace(frame_times=[40.0 * t for t in range(n)], frame_indices=list(range(n)), height=h, width=w, cc_group_ages=ages,
                           cc_group_images=images, cc_group_boundaries=bounds)
    kfs, times = KeyframeExtractor.GenerateFromST3DForIntervals(st3d, segs, False)
    check(kfs, times, content, ref_times)
