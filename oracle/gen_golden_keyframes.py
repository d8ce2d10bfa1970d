"""oracle/gen_golden_keyframes.py -- regenerates tests/golden/keyframes.npz (build container only: needs /root/reference).

Runs the UNMODIFIED reference: stage 02 + stage 03 on the seeded mask videos of tests/golden/cc_stability.npz (as
oracle/gen_golden_grouping.py does), wraps the stage-03 results in the reference's SpaceTimeStruct and calls
KeyframeExtractor.GenerateFromST3DForIntervals (R/AccessMath/preprocessing/content/keyframe_extractor.py:12-150) on a few frame
intervals.  Stored per run: the stage-03 inputs of the call (group ages / boundaries / images) and its outputs (key-frames, times).
The reference uses `np.bool` (:72), which NumPy >= 1.24 no longer has; the alias is restored for the call -- nothing else is touched.
   python oracle/gen_golden_keyframes.py
"""
import contextlib
import io
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle.gen_golden import GOLD, import_reference      # noqa: E402
from oracle.gen_golden_grouping import RUNS               # noqa: E402


def intervals(n_frames):
    a, b = n_frames // 3, 2 * n_frames // 3
    return [(0, a), (a + 1, b), (b + 1, n_frames - 1), (a // 2, a // 2 + 2), (0, n_frames - 1)]


def pack_st3d(ages, images, bounds):
    """group dictionaries -> arrays (shared with the tests, which rebuild the dictionaries from them)."""
    out = {}
    out["group_ages"] = np.array([(g, a) for g in ages for a in ages[g]], dtype=np.int64).reshape(-1, 2)        # dictionary order kept
    out["group_bounds"] = np.array([(g,) + tuple(int(v) for v in bounds[g]) for g in images], dtype=np.int64).reshape(-1, 5)
    out["image_shapes"] = np.array([(g, s) + im.shape for g in images for s, im in enumerate(images[g])], dtype=np.int64).reshape(-1, 4)
    flat = [im.ravel() for g in images for im in images[g]]
    out["image_bits"] = np.packbits(np.concatenate(flat) > 0) if flat else np.zeros(0, np.uint8)
    return out


def unpack_st3d(z, prefix):
    ages, bounds, images = {}, {}, {}
    for g, a in z[prefix + "group_ages"]:
        ages.setdefault(int(g), []).append(int(a))
    for row in z[prefix + "group_bounds"]:
        bounds[int(row[0])] = tuple(int(v) for v in row[1:])
    bits = np.unpackbits(z[prefix + "image_bits"])
    pos = 0
    for g, s, h, w in z[prefix + "image_shapes"]:
        n = int(h) * int(w)
        images.setdefault(int(g), []).append(np.ascontiguousarray(bits[pos:pos + n].reshape(int(h), int(w)).astype(np.uint8) * 255))
        pos += n
    return ages, images, bounds


def main():
    Labeler, CCStabilityEstimator, *_ = import_reference()
    np.bool = bool                                                        # removed from NumPy; keyframe_extractor.py:72 still uses it
    from AccessMath.data.space_time_struct import SpaceTimeStruct
    from AccessMath.preprocessing.content.keyframe_extractor import KeyframeExtractor
    z = np.load(os.path.join(GOLD, "cc_stability.npz"))
    out = {}
    for name, (split_gap, min_times, t_window, g_recall, img_t) in RUNS.items():
        n, h, w = (int(v) for v in z[name + "_shape"])
        masks = np.unpackbits(z[name + "_masks"], axis=-1)[:, :, :w].astype(np.uint8) * 255
        r, p, gap = z[name + "_params"]
        est = CCStabilityEstimator(w, h, float(r), float(p), int(gap), False)
        with contextlib.redirect_stdout(io.StringIO()), np.errstate(all="ignore"):
            for m in masks:
                est.add_frame(m, True)
            est.split_stable_cc_by_gaps(split_gap, min_times)
            stable = est.get_stable_cc_idxs(min_times)
            time_ov, _, all_ov = est.compute_overlapping_stable_cc(stable, t_window)
            groups, gidx = est.compute_groups(stable, time_ov, g_recall, None, None)
            ages, gpf = est.compute_groups_temporal_information(groups)
            images, bounds = est.compute_group_images(groups, ages, img_t)
            frame_times = [40.0 * t for t in range(n)]
            st3d = SpaceTimeStruct(frame_times, list(range(n)), h, w, ages, images, bounds)
            segs = intervals(n)
            kfs, times = KeyframeExtractor.GenerateFromST3DForIntervals(st3d, segs, False)
        for k, v in pack_st3d(ages, images, bounds).items():
            out[name + "/" + k] = v
        out[name + "/shape"] = np.array([n, h, w])
        out[name + "/segments"] = np.array(segs, dtype=np.int64)
        out[name + "/keyframes"] = np.packbits(np.stack(kfs)[:, :, :, 0] == 0, axis=-1)                # content pixels (value 0)
        out[name + "/keyframes_gray"] = np.array(int(all((k[:, :, 0] == k[:, :, 1]).all() and (k[:, :, 0] == k[:, :, 2]).all() and
                                                         set(np.unique(k)) <= {0, 255} for k in kfs)))
        out[name + "/times"] = np.array([(s,) + tuple(float(v) for v in row) for s, rows in enumerate(times) for row in rows],
                                        dtype=np.float64).reshape(-1, 6)
        print(name, "groups", len(ages), "keyframes", len(kfs), "time rows", len(out[name + "/times"]), "gray", int(out[name + "/keyframes_gray"]))
    np.savez_compressed(os.path.join(GOLD, "keyframes.npz"), **out)
    print("wrote", os.path.getsize(os.path.join(GOLD, "keyframes.npz")), "bytes")


if __name__ == "__main__":
    main()
