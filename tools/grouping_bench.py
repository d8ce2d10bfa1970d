"""Stage-03 (CC grouping) timing on dense 1080p masks: the CUDA drop-in's estimator methods, per-method wall-clock ms (device
work is synchronised inside each method by its read-back) and CUDA-event time of the C-ABI calls.
   python tools/grouping_bench.py [--frames 48]          (the CPU-oracle timing of the same workload: oracle/time_grouping_oracle.py)"""
import argparse
import contextlib
import io
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def stage03(est, timings):
    def run(name, fn, *a):
        t0 = time.perf_counter()
        with contextlib.redirect_stdout(io.StringIO()), np.errstate(all="ignore"):
            r = fn(*a)
        timings[name] = round(1000.0 * (time.perf_counter() - t0), 2)
        return r
    run("rebuilt_binary_images", est.rebuilt_binary_images)
    run("split_stable_cc_by_gaps", est.split_stable_cc_by_gaps, 85, 3)
    stable = run("get_stable_cc_idxs", est.get_stable_cc_idxs, 3)
    time_ov, total, all_ov = run("compute_overlapping_stable_cc", est.compute_overlapping_stable_cc, stable, 5)
    groups, gidx = run("compute_groups", est.compute_groups, stable, time_ov, 0.5, None, None)
    ages, gpf = run("compute_groups_temporal_information", est.compute_groups_temporal_information, groups)
    run("compute_conflicting_groups", est.compute_conflicting_groups, stable, all_ov, len(groups), gidx)
    images, bounds = run("compute_group_images", est.compute_group_images, groups, ages, 0.5)
    clean = run("frames_from_groups", est.frames_from_groups, groups, bounds, gpf, ages, images, None, 3, True)
    return {"stable": len(stable), "pairs": sum(len(x) for x in all_ov) // 2, "groups": len(groups),
            "segments": sum(len(v) for v in images.values()), "frames": len(clean)}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--frames", type=int, default=48)
    ap.add_argument("--gc-report", action="store_true", help="also report the time spent inside Python's cyclic garbage collector")
    ap.add_argument("--no-oracle", action="store_true", help="(kept for old command lines; the CPU-oracle timing lives in oracle/time_grouping_oracle.py)")
    args = ap.parse_args()
    import torch
    from lecturemath_b200 import synth
    from lecturemath_b200.cc_stability_estimator import CCStabilityEstimator
    h, w = 1080, 1920
    masks = np.stack(list(synth.glyph_masks(args.frames, h, w, seed=0)))
    est = CCStabilityEstimator(w, h, 0.85, 0.85, 85)
    est.add_frames(masks)
    torch.cuda.synchronize()
    t_gpu = {}
    stage03(est, {})                                                       # warm-up on a copy of the state would change it: re-run below
    est = CCStabilityEstimator(w, h, 0.85, 0.85, 85)
    est.add_frames(masks)
    torch.cuda.synchronize()
    est.device_ms = {}
    gc_ms = [0.0, 0]
    if args.gc_report:
        import gc
        t_gc = [0.0]

        def on_gc(phase, info):
            if phase == "start":
                t_gc[0] = time.perf_counter()
            else:
                gc_ms[0] += 1000.0 * (time.perf_counter() - t_gc[0]); gc_ms[1] += 1
        gc.callbacks.append(on_gc)
    info = stage03(est, t_gpu)
    if args.gc_report:
        gc.callbacks.remove(on_gc)
        info["gc_ms"], info["gc_runs"] = round(gc_ms[0], 1), gc_ms[1]
    line = {"workload": "stage 03 on %d dense 1080p glyph-mask frames" % args.frames, **info, "gpu_ms": t_gpu,
            "gpu_total_ms": round(sum(t_gpu.values()), 1),
            "device_kernel_ms": {k: round(v, 3) for k, v in est.device_ms.items()}}
    print(json.dumps(line))


if __name__ == "__main__":
    main()
