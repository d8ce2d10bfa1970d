"""CPU: the order-dependent bookkeeping of stage 03 that stays on the host (lecturemath_b200/cc_grouping.py: split by gaps, stable
list, group merging / numbering, ages, conflicts) against the outputs of the unmodified reference (tests/golden/cc_grouping.npz).
The pixel work these methods consume (pair overlaps) is injected from the oracle here; the CUDA kernels that produce it on the
product path are covered by tests/test_grouping_gpu.py."""
import numpy as np
import pytest

from lecturemath_b200.cc_grouping import GroupingMixin
from lecturemath_b200.connected_component import ConnectedComponent
from oracle import cc_oracle as CO
from oracle.gen_golden_grouping import RUNS
from oracle.grouping_oracle import GroupingOracle
from tests.conftest import unpack_masks


class HostOnly(GroupingMixin):
    """The drop-in's Python-visible state without any device handle."""

    def __init__(self, stab):
        self.width, self.height = stab.width, stab.height
        conv = {}

        def cc_of(o):
            if id(o) not in conv:
                conv[id(o)] = ConnectedComponent(o.cc_id, np.int32(o.min_x), np.int32(o.max_x), np.int32(o.min_y), np.int32(o.max_y),
                                                 np.int32(o.size), img=o.img)
            return conv[id(o)]
        self.unique_cc_objects = [cc_of(o) for o in stab.unique_cc_objects]
        self.unique_cc_frames = [list(f) for f in stab.unique_cc_frames]
        self.cc_idx_per_frame = [[(u, cc_of(o)) for u, o in row] for row in stab.cc_idx_per_frame]


@pytest.mark.parametrize("name", sorted(RUNS))
def test_host_bookkeeping_matches_reference(golden, name):
    zs, zg = golden("cc_stability.npz"), golden("cc_grouping.npz")
    masks = unpack_masks(zs, name)
    r, p, gap = zs[name + "_params"]
    stab = CO.StabilityOracle(masks.shape[2], masks.shape[1], float(r), float(p), int(gap))
    for m in masks:
        stab.add_frame(m)
    split_gap, min_times, t_window, g_recall, _ = zg[name + "/params"]
    host, orc = HostOnly(stab), GroupingOracle(stab)
    assert host.split_stable_cc_by_gaps(int(split_gap), int(min_times)) == int(zg[name + "/split_count"])
    orc.split_stable_cc_by_gaps(int(split_gap), int(min_times))
    got = np.array([(u, t, lab) for u, lst in enumerate(host.unique_cc_frames) for t, lab in lst], dtype=np.int64).reshape(-1, 3)
    np.testing.assert_array_equal(got, zg[name + "/uframes"])
    got = np.array([(t, u, cc.cc_id + 1) for t, fr in enumerate(host.cc_idx_per_frame) for u, cc in fr], dtype=np.int64).reshape(-1, 3)
    np.testing.assert_array_equal(got, zg[name + "/per_frame"])
    stable = host.get_stable_cc_idxs(int(min_times))
    np.testing.assert_array_equal(np.array(stable), zg[name + "/stable"])
    time_ov, _, all_ov = orc.compute_overlapping_stable_cc(stable, int(t_window))      # injected pixel work
    groups, gidx = host.compute_groups(stable, time_ov, float(g_recall), None, None)
    np.testing.assert_array_equal(np.array([(g, u) for g, grp in enumerate(groups) for u in grp], dtype=np.int64).reshape(-1, 2), zg[name + "/groups"])
    np.testing.assert_array_equal(np.array(sorted(gidx.items()), dtype=np.int64).reshape(-1, 2), zg[name + "/group_idx_per_cc"])
    ages, gpf = host.compute_groups_temporal_information(groups)
    np.testing.assert_array_equal(np.array([(g, a) for g in sorted(ages) for a in ages[g]], dtype=np.int64).reshape(-1, 2), zg[name + "/group_ages"])
    np.testing.assert_array_equal(np.array([(t, g) for t, lst in enumerate(gpf) for g in lst], dtype=np.int64).reshape(-1, 2),
                                  zg[name + "/groups_per_frame"])
    conflicts = host.compute_conflicting_groups(stable, all_ov, len(groups), gidx)
    got = np.array([(g1, g2, d["matched"], d["unmatched"], d["area_union"], d["area_intersection"])
                    for g1 in sorted(conflicts) for g2, d in conflicts[g1].items()], dtype=np.int64).reshape(-1, 6)
    np.testing.assert_array_equal(got, zg[name + "/conflicts"])
    assert host.get_temporal_index() == [[u for u, _ in row] for row in host.cc_idx_per_frame]


def test_pixel_methods_fail_loudly_without_a_device():
    """No CPU fallback: without a CUDA device the pixel-work methods raise (they never route through the oracle)."""
    from lecturemath_b200 import _lib
    if _lib.load().am_device_count() > 0:
        pytest.skip("a CUDA device is present")
    stab = CO.StabilityOracle(64, 48, 0.85, 0.85, 5)
    m = np.zeros((48, 64), np.uint8)
    m[10:20, 10:30] = 255
    for _ in range(3):
        stab.add_frame(m)
    host = HostOnly(stab)
    with pytest.raises(_lib.AccessMathB200Error):
        host.rebuilt_binary_images()
    with pytest.raises(_lib.AccessMathB200Error):
        host.compute_group_images([[0]], {0: [0, 2]}, 0.5)


def test_frame_segment_items_equals_the_reference_cursor():
    """frames_from_groups picks, per frame and alive group, the age segment the reference's per-group cursor stands on
    (cc_stability_estimator.py:655-660); the drop-in finds all of them with one searchsorted -- same rows, same order."""
    import random
    from lecturemath_b200.cc_grouping import frame_segment_items
    rng = random.Random(7)
    for _ in range(150):
        n_frames, n_groups = rng.randint(1, 40), rng.randint(1, 30)
        ages, per_frame, seg_index = {}, [[] for _ in range(n_frames)], {}
        for g in range(n_groups):
            marks = sorted(rng.sample(range(n_frames + 5), min(rng.randint(2, 6), n_frames + 5)))
            if rng.random() < 0.2 or len(marks) < 2:
                continue
            ages[g] = marks
            for s in range(len(marks) - 1):
                seg_index[(g, s)] = len(seg_index)
            for t in range(marks[0], min(marks[-1] + 1, n_frames)):
                per_frame[t].append(g)
        cursor, want = [0] * n_groups, []
        for t, present in enumerate(per_frame):
            for g in present:
                while ages[g][cursor[g] + 1] < t:
                    cursor[g] += 1
                want.append((t, seg_index[(g, cursor[g])]))
        np.testing.assert_array_equal(frame_segment_items(per_frame, ages, seg_index, n_groups), np.array(want, dtype=np.int32).reshape(-1, 2))
