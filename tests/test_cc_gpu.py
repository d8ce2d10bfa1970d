"""GPU parity tests for the CC stage, through the C ABI (libaccessmath_b200.so), against the CPU oracle and the
reference's captured outputs.  Bar: bit-exact (labels, statistics, crops, match lists, tempo_count)."""
import ctypes

import numpy as np
import pytest
import torch

from oracle import cc_oracle as O
from tests.conftest import unpack_masks

pytestmark = pytest.mark.gpu

CASES = ["blob_96x128", "blob_67x121", "glyph_180x320", "empty_40x70", "full_33x65", "noise_50x97"]


def _engine(w, h, b=1, min_pixels=20):
    from lecturemath_b200.cc_engine import CCEngine
    return CCEngine(w, h, b, min_pixels)


def _label_frames(masks, want_labels=True, min_pixels=20):
    b, h, w = masks.shape
    eng = _engine(w, h, b, min_pixels)
    bits = eng.pack(torch.from_numpy(masks).cuda())
    labels = eng.label(bits, want_labels=want_labels)
    return eng, bits, labels


@pytest.mark.parametrize("case", CASES)
def test_label_stats_crops_vs_reference_golden(golden, case):
    from lecturemath_b200.connected_component import unpack_crop
    z = golden("cc_label_stats.npz")
    mask = z[case + "_mask"]
    eng, bits, labels = _label_frames(mask[None])
    assert int(eng.counts[0, 1]) == int(z[case + "_n"])
    np.testing.assert_array_equal(labels[0].cpu().numpy(), z[case + "_labels"])
    np.testing.assert_array_equal(eng.unpack(bits)[0].cpu().numpy(), (mask != 0).astype(np.uint8) * 255)
    rows, crops = eng.kept_rows(0), eng.crops(0)
    np.testing.assert_array_equal(rows[:, 1:7].astype(np.int64), z[case + "_table"])
    got = [unpack_crop(crops[r[7]:r[7] + ((r[3] >> 5) - (r[2] >> 5) + 1) * (r[5] - r[4] + 1)], r[2], r[3], r[4], r[5]).ravel() for r in rows]
    got = np.concatenate(got) if got else np.zeros(0, np.uint8)
    np.testing.assert_array_equal(got, z[case + "_crops"])


def test_label_table_equals_age_boundaries_all_labels():
    rng = np.random.default_rng(3)
    masks = (rng.random((3, 75, 203)) < np.array([0.3, 0.55, 0.7])[:, None, None]).astype(np.uint8)
    eng, _, labels = _label_frames(masks)
    for f in range(3):
        lab_o, n = O.label4(masks[f])
        assert int(eng.counts[f, 1]) == n
        np.testing.assert_array_equal(labels[f].cpu().numpy(), lab_o)
        ref = O.age_boundaries(lab_o, np.zeros(lab_o.shape, np.float32), n)
        got = eng.label_table(f)
        for a, b in zip(got, ref[:5]):
            np.testing.assert_array_equal(a, b)


def test_edge_shapes():
    rng = np.random.default_rng(4)
    for h, w, d in [(1, 1, 1.0), (1, 33, 0.7), (40, 1, 0.7), (2, 64, 1.0), (31, 32, 0.5), (9, 257, 0.6), (64, 96, 0.0)]:
        m = (rng.random((h, w)) < d).astype(np.uint8)
        eng, _, labels = _label_frames(m[None], min_pixels=1)
        lab_o, n = O.label4(m)
        assert int(eng.counts[0, 1]) == n
        np.testing.assert_array_equal(labels[0].cpu().numpy(), lab_o)
        assert int(eng.counts[0, 2]) == n            # min_pixels=1 keeps everything


def test_serpentine_and_spiral_components():
    # long dependency chains for the union-find: a serpentine that is ONE component, and comb shapes
    h, w = 129, 257
    m = np.zeros((h, w), np.uint8)
    m[::2, :] = 1
    for i, y in enumerate(range(1, h, 2)):
        m[y, -1 if i % 2 == 0 else 0] = 1
    comb = np.zeros((h, w), np.uint8)
    comb[-1, :] = 1
    comb[:, ::2] = 1
    for mask in (m, comb, comb[::-1].copy()):
        eng, _, labels = _label_frames(mask[None], min_pixels=1)
        lab_o, n = O.label4(mask)
        assert n == 1 and int(eng.counts[0, 1]) == 1
        np.testing.assert_array_equal(labels[0].cpu().numpy(), lab_o)


def test_legacy_cc_age_boundaries_host_abi():
    from lecturemath_b200 import _lib
    lib = _lib.lib()
    rng = np.random.default_rng(5)
    m = (rng.random((83, 140)) < 0.5).astype(np.uint8)
    lab, n = O.label4(m)
    ages = (rng.random(m.shape) * 99).astype(np.float32)
    ref = O.age_boundaries(lab, ages, n)
    outs = [np.zeros(n, np.int32) for _ in range(5)] + [np.zeros(n, np.float32)]
    p = lambda a: a.ctypes.data_as(ctypes.c_void_p)
    assert lib.CC_AgeBoundaries(p(lab), p(ages), 140, 83, n, *[p(o) for o in outs]) == 0
    for a, b in zip(outs, ref):
        np.testing.assert_array_equal(a, b)
    # a label count larger than what the image holds: untouched rows keep the init values (:364-374)
    outs = [np.zeros(n + 3, np.int32) for _ in range(5)] + [np.zeros(n + 3, np.float32)]
    assert lib.CC_AgeBoundaries(p(lab), p(ages), 140, 83, n + 3, *[p(o) for o in outs]) == 0
    assert outs[4][-1] == 0 and outs[5][-1] == -1.0 and outs[0][-1] == 83 and outs[2][-1] == 140


@pytest.mark.parametrize("run", ["blobs_gap6", "blobs_gap85", "glyphs", "loose"])
@pytest.mark.parametrize("batch", [1, 7])
def test_stability_vs_reference_golden(golden, run, batch):
    from lecturemath_b200.cc_stability_estimator import CCStabilityEstimator
    z = golden("cc_stability.npz")
    masks = unpack_masks(z, run)
    r, p, gap = z[run + "_params"]
    est = CCStabilityEstimator(masks.shape[2], masks.shape[1], float(r), float(p), int(gap), False, max_batch=batch)
    if batch == 1:
        for m in masks:
            est.add_frame(m, True)
    else:
        est.add_frames(masks)
    per_frame = np.array([(t, u, cc.cc_id + 1, cc.min_x, cc.max_x, cc.min_y, cc.max_y, cc.size)
                          for t, fr in enumerate(est.cc_idx_per_frame) for u, cc in fr], dtype=np.int64).reshape(-1, 8)
    np.testing.assert_array_equal(per_frame, z[run + "_per_frame"])
    uf = np.array([(u, t, l) for u, lst in enumerate(est.unique_cc_frames) for t, l in lst], dtype=np.int64).reshape(-1, 3)
    np.testing.assert_array_equal(uf, z[run + "_uframes"])
    assert est.tempo_count == int(z[run + "_tempo"])
    uq = np.array([(c.cc_id + 1, c.min_x, c.max_x, c.min_y, c.max_y, c.size) for c in est.unique_cc_objects], dtype=np.int64).reshape(-1, 6)
    np.testing.assert_array_equal(uq, z[run + "_uniq"])
    assert est.get_raw_cc_count() == len(per_frame)


def test_labeler_dropin_vs_oracle_with_ages():
    from lecturemath_b200.labeler import Labeler
    from lecturemath_b200 import synth
    mask = next(iter(synth.glyph_masks(1, 200, 352, seed=5, occluder_w=0)))
    ages = np.random.default_rng(0).random(mask.shape).astype(np.float32) * 10
    got = Labeler.extractSpatioTemporalContent(mask, ages)
    ref, _, _ = O.extract_components(mask, ages)
    assert len(got) == len(ref) > 50
    for a, b in zip(got, ref):
        assert (a.cc_id, a.min_x, a.max_x, a.min_y, a.max_y, a.size) == (b.cc_id, b.min_x, b.max_x, b.min_y, b.max_y, b.size)
        assert a.start_time == b.start_time
        np.testing.assert_array_equal(a.img, b.img)
    assert Labeler.extractConnectedComponents(np.zeros((30, 50), np.uint8)) == []


def test_full_size_1080p_dense_masks_properties_and_oracle():
    """BASELINE config 4 shape: 1080p dense glyph masks (>5k CCs/frame); oracle comparison on 3 frames plus
    size-independent properties: counts sum to ink pixels, bbox contains crop popcount == size."""
    from lecturemath_b200 import synth
    masks = np.stack(list(synth.glyph_masks(3, 1080, 1920, seed=0)))
    eng, bits, labels = _label_frames(masks)
    for f in range(3):
        lab_o, n = O.label4(masks[f])
        assert int(eng.counts[f, 1]) == n
        np.testing.assert_array_equal(labels[f].cpu().numpy(), lab_o)
        t = eng.label_table(f)
        assert int(t[4].sum()) == int((masks[f] != 0).sum())
        rows, crops = eng.kept_rows(f), eng.crops(f)
        assert len(rows) > 4000
        pop = np.array([int(np.unpackbits(crops[r[7]:r[7] + ((r[3] >> 5) - (r[2] >> 5) + 1) * (r[5] - r[4] + 1)].view(np.uint8)).sum()) for r in rows[:500]])
        np.testing.assert_array_equal(pop, rows[:500, 6])


def test_estimator_state_export_import_roundtrip(golden):
    """Frame-shard hand-off on one GPU: run frames [0,k) in estimator A, export the active set, import into a
    fresh estimator B, run [k,n): results must equal the single-estimator run."""
    from lecturemath_b200.cc_engine import CCEngine, Estimator
    z = golden("cc_stability.npz")
    masks = unpack_masks(z, "blobs_gap6")
    n, h, w = masks.shape
    r, p, gap = z["blobs_gap6_params"]
    ref = z["blobs_gap6_per_frame"]
    dev = torch.from_numpy(masks).cuda()
    for k in (1, 23, 40):
        eng = CCEngine(w, h, n)
        eng.label(eng.pack(dev), sync=False)
        a = Estimator(w, h, float(r), float(p), int(gap))
        a.add_frames(eng, 0, k)
        header, meta, crops = a.export_state()
        b = Estimator(w, h, float(r), float(p), int(gap))
        b.import_state(header, meta, crops)
        b.add_frames(eng, k, n - k)
        eng.read_counts()
        rows, offs = eng.packed_rows(n)
        rows, offs = rows.cpu().numpy(), offs.cpu().numpy()
        got = np.array([(t,) + tuple(int(v) for v in row[:7]) for t in range(n) for row in rows[offs[t]:offs[t + 1]]], dtype=np.int64).reshape(-1, 8)
        np.testing.assert_array_equal(got, ref)
        assert b.state()["tempo_count"] == int(z["blobs_gap6_tempo"])


def _chalk_masks_4k(n):
    """Sparse 3840x2160 masks: the ink of the synthetic chalkboard video (strokes 230 on a 40 background, BASELINE configs[4])."""
    from lecturemath_b200 import synth
    return np.stack([(fr.max(axis=2) > 128).astype(np.uint8) * 255
                     for fr in synth.whiteboard_frames(n, 2160, 3840, seed=21, chalk=True, strokes_per_frame=40)])


@pytest.mark.parametrize("kind", ["chalk", "dense"])
def test_4k_label_rows_crops_and_match_vs_oracle(kind):
    """BASELINE configs[4]: the CC stage at 3840x2160 (am_cc_create picks R = 8 rows per strip there, a configuration the
    1080p cases never reach).  Labels, the label table, kept rows, crops and an 8-frame temporal match: bit-exact vs the oracle."""
    from lecturemath_b200 import synth
    from lecturemath_b200.cc_engine import CCEngine, Estimator
    from lecturemath_b200.connected_component import unpack_crop
    n, h, w = 8, 2160, 3840
    masks = _chalk_masks_4k(n) if kind == "chalk" else np.stack(list(synth.glyph_masks(n, h, w, seed=2)))
    eng = CCEngine(w, h, n)
    bits = eng.pack(torch.from_numpy(masks).cuda())
    labels = eng.label(bits, want_labels=True)
    est_o = O.StabilityOracle(w, h, 0.85, 0.85, 85)
    est = Estimator(w, h, 0.85, 0.85, 85)
    est.add_frames(eng, 0, n)
    rows_d, offs_d = eng.packed_rows(n)
    rows_d, offs_d = rows_d.cpu().numpy(), offs_d.cpu().numpy()
    for f in range(n):
        lab_o, n_lab = O.label4(masks[f])
        assert int(eng.counts[f, 1]) == n_lab
        if f < 2:                                                       # the 33 MB label image: two frames are enough
            np.testing.assert_array_equal(labels[f].cpu().numpy(), lab_o)
            ref_t = O.age_boundaries(lab_o, np.zeros(lab_o.shape, np.float32), n_lab)
            for a, b in zip(eng.label_table(f), ref_t[:5]):
                np.testing.assert_array_equal(a, b)
        est_o.add_frame(masks[f])
        ref = np.array(est_o.frame_table(f), dtype=np.int64).reshape(-1, 7)
        np.testing.assert_array_equal(rows_d[offs_d[f]:offs_d[f + 1], :7].astype(np.int64), ref)
        if f in (0, n - 1):                                             # crops of every kept CC against the oracle's images
            rows, crops = eng.kept_rows(f), eng.crops(f)
            comps, _, _ = O.extract_components(masks[f])
            assert len(rows) == len(comps)
            for r, c in zip(rows, comps):
                got = unpack_crop(crops[r[7]:r[7] + ((r[3] >> 5) - (r[2] >> 5) + 1) * (r[5] - r[4] + 1)], r[2], r[3], r[4], r[5])
                np.testing.assert_array_equal(got, c.img)
    st = est.state()
    assert st["tempo_count"] == est_o.tempo_count and st["n_unique"] == len(est_o.unique_cc_objects)
    assert (len(est_o.unique_cc_objects) > 15000) if kind == "dense" else (len(est_o.unique_cc_objects) > 50)


def test_capacity_overflow_grows_and_retries(golden):
    """The reference has no capacity limits.  A synchronous label() whose kept-CC / crop capacities overflow doubles them and labels
    the same masks again (CCEngine._enlarge) instead of failing; the asynchronous pipelines still fail loudly (their temporal state
    has already consumed the truncated tables)."""
    from lecturemath_b200.cc_engine import CCEngine
    from lecturemath_b200._lib import AccessMathB200Error
    z = golden("cc_label_stats.npz")
    mask = z["glyph_180x320_mask"]
    eng = CCEngine(320, 180, 1, max_kept=64, crop_words=1024)
    bits = eng.pack(torch.from_numpy(mask[None]).cuda())
    eng.label(bits, want_labels=False)                                 # sync=True: overflows, grows, relabels
    assert eng._grow >= 1
    rows = eng.kept_rows(0)
    np.testing.assert_array_equal(rows[:, 1:7].astype(np.int64), z["glyph_180x320_table"])
    tiny = CCEngine(320, 180, 1, max_kept=64, crop_words=1024)
    tiny.label(bits, want_labels=False, sync=False)
    with pytest.raises(AccessMathB200Error):
        tiny.read_counts()
