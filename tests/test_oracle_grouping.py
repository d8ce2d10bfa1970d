"""CPU tests: the stage-03 grouping oracle (oracle/grouping_oracle.py) against the outputs of the unmodified reference
(tests/golden/cc_grouping.npz): every list, table and image the nine estimator methods produce, in the reference's order."""
import numpy as np
import pytest

from oracle import cc_oracle as CO
from oracle.gen_golden_grouping import RUNS, run_stage03
from oracle.grouping_oracle import GroupingOracle
from tests.conftest import unpack_masks

EXACT_KEYS = ["rebuilt", "rebuilt_is_binary", "split_count", "n_objects", "uframes", "per_frame", "stable", "time_ov_idx", "time_ov_rp",
              "total_intersections", "all_ov", "groups", "group_idx_per_cc", "group_ages", "groups_per_frame", "conflicts",
              "group_bounds", "group_image_shapes", "group_image_bits", "group_images_are_binary", "clean"]


def compare_with_golden(z, name, res):
    for k in EXACT_KEYS:
        ref, got = z[name + "/" + k], res[k]
        assert ref.shape == got.shape, "%s/%s: shape %s != %s" % (name, k, got.shape, ref.shape)
        assert np.array_equal(ref, got), "%s/%s differs from the reference" % (name, k)      # fp64 recall/precision included: bit-exact


@pytest.mark.parametrize("name", sorted(RUNS))
def test_grouping_oracle_matches_reference(golden, name):
    zs, zg = golden("cc_stability.npz"), golden("cc_grouping.npz")
    masks = unpack_masks(zs, name)
    r, p, gap = zs[name + "_params"]
    stab = CO.StabilityOracle(masks.shape[2], masks.shape[1], float(r), float(p), int(gap))
    for m in masks:
        stab.add_frame(m)
    split_gap, min_times, t_window, g_recall, img_t = zg[name + "/params"]
    res = run_stage03(GroupingOracle(stab), int(split_gap), int(min_times), int(t_window), float(g_recall), float(img_t))
    compare_with_golden(zg, name, res)
