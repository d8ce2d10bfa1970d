"""CPU: the CC oracle against the reference's captured outputs (tests/golden, oracle/gen_golden.py),
against SciPy (the reference's labeler) and against the reference's own compiled C (oracle/_ref)."""
import numpy as np
import pytest
import scipy.ndimage

from oracle import cc_oracle as O
from tests.conftest import unpack_masks

CASES = ["blob_96x128", "blob_67x121", "glyph_180x320", "empty_40x70", "full_33x65", "noise_50x97"]


def test_known_answer_masks(golden):
    z = golden("cc_known_answer.npz")
    for name in ("diag", "ushape"):
        lab, n = O.label4(z[name + "_mask"])
        assert n == int(z[name + "_n"])
        np.testing.assert_array_equal(lab, z[name + "_labels"])
    # diagonal pixels are separate components under 4-connectivity
    assert int(z["diag_n"]) >= 6


@pytest.mark.parametrize("case", CASES)
def test_label_stats_crops_match_reference(golden, case):
    z = golden("cc_label_stats.npz")
    mask, ages = z[case + "_mask"], z[case + "_ages"]
    lab, n = O.label4(mask)
    assert n == int(z[case + "_n"])
    np.testing.assert_array_equal(lab, z[case + "_labels"])
    comps, _, _ = O.extract_components(mask, ages)
    tab = np.array([(c.cc_id + 1, c.min_x, c.max_x, c.min_y, c.max_y, c.size) for c in comps], dtype=np.int64).reshape(-1, 6)
    np.testing.assert_array_equal(tab, z[case + "_table"])
    np.testing.assert_array_equal(np.array([c.start_time for c in comps], dtype=np.float32), z[case + "_minage"])
    crops = np.concatenate([c.img.ravel() for c in comps]) if comps else np.zeros(0, np.uint8)
    np.testing.assert_array_equal(crops, z[case + "_crops"])


def test_label4_equals_scipy_random():
    rng = np.random.default_rng(0)
    for h, w, d in [(1, 1, 1.0), (1, 40, 0.5), (37, 1, 0.6), (64, 64, 0.5), (50, 131, 0.62), (33, 200, 0.3)]:
        m = (rng.random((h, w)) < d).astype(np.uint8)
        ref, n_ref = scipy.ndimage.label(m)
        lab, n = O.label4(m)
        assert n == n_ref
        np.testing.assert_array_equal(lab, ref.astype(np.int32))


def test_age_boundaries_equals_reference_c():
    if O.ref_lib() is None:
        pytest.skip("oracle/_ref not built (reference source absent and no prebuilt copy)")
    rng = np.random.default_rng(1)
    m = (rng.random((70, 93)) < 0.5).astype(np.uint8)
    lab, n = O.label4(m)
    ages = (rng.random(m.shape) * 50).astype(np.float32)
    for a, b in zip(O.age_boundaries(lab, ages, n), O.age_boundaries_ref(lab, ages, n)):
        np.testing.assert_array_equal(a, b)


@pytest.mark.parametrize("run", ["blobs_gap6", "blobs_gap85", "glyphs", "loose"])
def test_stability_oracle_matches_reference(golden, run):
    z = golden("cc_stability.npz")
    masks = unpack_masks(z, run)
    r, p, gap = z[run + "_params"]
    est = O.StabilityOracle(masks.shape[2], masks.shape[1], float(r), float(p), int(gap))
    for m in masks:
        est.add_frame(m)
    per_frame = np.array([(t,) + row for t in range(len(masks)) for row in est.frame_table(t)], dtype=np.int64).reshape(-1, 8)
    np.testing.assert_array_equal(per_frame, z[run + "_per_frame"])
    uf = np.array([(u, t, l) for u, lst in enumerate(est.unique_cc_frames) for t, l in lst], dtype=np.int64).reshape(-1, 3)
    np.testing.assert_array_equal(uf, z[run + "_uframes"])
    assert est.finish_processing() == int(z[run + "_tempo"])
    uq = np.array([(c.cc_id + 1, c.min_x, c.max_x, c.min_y, c.max_y, c.size) for c in est.unique_cc_objects], dtype=np.int64).reshape(-1, 6)
    np.testing.assert_array_equal(uq, z[run + "_uniq"])
