"""oracle/gen_golden_grouping.py -- regenerates tests/golden/cc_grouping.npz (build container only: needs /root/reference).

Runs the UNMODIFIED reference through stage 02 (CCStabilityEstimator.add_frame on the seeded mask videos already stored in
tests/golden/cc_stability.npz) and then through every estimator method stage 03 calls
(R/pre_ST3D_v3.0_03_cc_grouping.py:41-101; R/AccessMath/preprocessing/content/cc_stability_estimator.py:166-681):
rebuilt_binary_images, split_stable_cc_by_gaps, get_stable_cc_idxs, compute_overlapping_stable_cc, compute_groups,
compute_groups_temporal_information, compute_conflicting_groups, compute_group_images, frames_from_groups.
Every result is flattened into integer / float64 arrays (list order preserved).
   python oracle/gen_golden_grouping.py
"""
import contextlib
import io
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle.gen_golden import GOLD, import_reference      # noqa: E402

# run name -> (split max_gap, stable min frames, temporal window, group min recall, image threshold)
RUNS = {"blobs_gap6": (6, 3, 5, 0.5, 0.5), "blobs_gap85": (10, 3, 5, 0.5, 0.5), "glyphs": (4, 3, 5, 0.5, 0.5),
        "loose": (2, 2, 3, 0.2, 0.35)}


def flatten_grouping(est, rebuilt, split_count, stable, time_ov, total_inter, all_ov, groups, gidx, ages, gpf, conflicts, images, bounds,
                     clean):
    """Reference-shaped results -> dict of arrays (shared by the generator and the tests' comparison helper)."""
    import cv2
    out = {}
    out["rebuilt"] = np.packbits(np.stack(rebuilt) > 0, axis=-1)
    out["rebuilt_is_binary"] = np.array(int(all(set(np.unique(r)) <= {0, 255} for r in rebuilt)))
    out["split_count"] = np.array(split_count)
    out["n_objects"] = np.array(len(est.unique_cc_objects))
    out["uframes"] = np.array([(u, t, lab) for u, lst in enumerate(est.unique_cc_frames) for t, lab in lst], dtype=np.int64).reshape(-1, 3)
    out["per_frame"] = np.array([(t, u, cc.cc_id + 1) for t, fr in enumerate(est.cc_idx_per_frame) for u, cc in fr], dtype=np.int64).reshape(-1, 3)
    out["stable"] = np.array(stable, dtype=np.int64)
    out["time_ov_idx"] = np.array([(a, b) for a, lst in enumerate(time_ov) for b, _, _ in lst], dtype=np.int64).reshape(-1, 2)
    out["time_ov_rp"] = np.array([(r, p) for lst in time_ov for _, r, p in lst], dtype=np.float64).reshape(-1, 2)
    out["total_intersections"] = np.array(total_inter)
    out["all_ov"] = np.array([(a, b, m, s2, s1) for a, lst in enumerate(all_ov) for b, m, s2, s1 in lst], dtype=np.int64).reshape(-1, 5)
    out["groups"] = np.array([(g, u) for g, grp in enumerate(groups) for u in grp], dtype=np.int64).reshape(-1, 2)
    out["group_idx_per_cc"] = np.array(sorted(gidx.items()), dtype=np.int64).reshape(-1, 2)
    out["group_ages"] = np.array([(g, a) for g in sorted(ages) for a in ages[g]], dtype=np.int64).reshape(-1, 2)
    out["groups_per_frame"] = np.array([(t, g) for t, lst in enumerate(gpf) for g in lst], dtype=np.int64).reshape(-1, 2)
    out["conflicts"] = np.array([(g1, g2, d["matched"], d["unmatched"], d["area_union"], d["area_intersection"])
                                 for g1 in sorted(conflicts) for g2, d in conflicts[g1].items()], dtype=np.int64).reshape(-1, 6)
    out["group_bounds"] = np.array([(g,) + tuple(int(v) for v in bounds[g]) for g in sorted(bounds)], dtype=np.int64).reshape(-1, 5)
    out["group_image_shapes"] = np.array([(g, s) + im.shape for g in sorted(images) for s, im in enumerate(images[g])], dtype=np.int64).reshape(-1, 4)
    flat = [im.ravel() for g in sorted(images) for im in images[g]]
    out["group_image_bits"] = np.packbits(np.concatenate(flat) > 0) if flat else np.zeros(0, np.uint8)
    out["group_images_are_binary"] = np.array(int(all(set(np.unique(f)) <= {0, 255} for f in flat)))
    out["clean"] = np.stack([cv2.imdecode(np.asarray(raw), cv2.IMREAD_GRAYSCALE) for raw in clean]) if clean else np.zeros((0, 1, 1), np.uint8)
    return out


def run_stage03(est, split_gap, min_times, t_window, g_recall, img_t):
    """The estimator-method sequence of pre_ST3D_v3.0_03_cc_grouping.py:41-101 (works on the reference class and on the drop-in)."""
    with contextlib.redirect_stdout(io.StringIO()), np.errstate(all="ignore"):
        rebuilt = est.rebuilt_binary_images()
        split_count = est.split_stable_cc_by_gaps(split_gap, min_times)
        stable = est.get_stable_cc_idxs(min_times)
        time_ov, total_inter, all_ov = est.compute_overlapping_stable_cc(stable, t_window)
        groups, gidx = est.compute_groups(stable, time_ov, g_recall, None, None)
        ages, gpf = est.compute_groups_temporal_information(groups)
        conflicts = est.compute_conflicting_groups(stable, all_ov, len(groups), gidx)
        images, bounds = est.compute_group_images(groups, ages, img_t)
        clean = est.frames_from_groups(groups, bounds, gpf, ages, images, None, min_times, True)
    return flatten_grouping(est, rebuilt, split_count, stable, time_ov, total_inter, all_ov, groups, gidx, ages, gpf, conflicts, images,
                            bounds, clean)


def main():
    Labeler, CCStabilityEstimator, FCN_LectureNet, FCN_LectureNet_Binarizer, Configuration = import_reference()
    z = np.load(os.path.join(GOLD, "cc_stability.npz"))
    out = {}
    for name, (split_gap, min_times, t_window, g_recall, img_t) in RUNS.items():
        n, h, w = (int(v) for v in z[name + "_shape"])
        masks = np.unpackbits(z[name + "_masks"], axis=-1)[:, :, :w].astype(np.uint8) * 255
        r, p, gap = z[name + "_params"]
        est = CCStabilityEstimator(w, h, float(r), float(p), int(gap), False)
        for m in masks:
            est.add_frame(m, True)
        res = run_stage03(est, split_gap, min_times, t_window, g_recall, img_t)
        for k, v in res.items():
            out[name + "/" + k] = v
        out[name + "/params"] = np.array([split_gap, min_times, t_window, g_recall, img_t], dtype=np.float64)
        print(name, "objects", int(res["n_objects"]), "split", int(res["split_count"]), "stable", len(res["stable"]), "pairs",
              len(res["all_ov"]) // 2, "time pairs", int(res["total_intersections"]), "groups", len(np.unique(res["groups"][:, 0])),
              "conflicts", len(res["conflicts"]), "segments", len(res["group_image_shapes"]), "clean values", np.unique(res["clean"])[-4:])
    np.savez_compressed(os.path.join(GOLD, "cc_grouping.npz"), **out)
    print("wrote", os.path.join(GOLD, "cc_grouping.npz"), os.path.getsize(os.path.join(GOLD, "cc_grouping.npz")), "bytes")


if __name__ == "__main__":
    main()
