"""Measures every feasible (S, Sy, NT, MT) packing / tiling of each FCN conv layer on the GPU and writes the winners to
lecturemath_b200/tuned_plans.json, which FCNPlan consults before its cycle model.

    python tools/autotune_fcn.py [--batch 8] [--hw 1080x1920] [--keep 12] [--iters 5] [--layers conv_up_block_1,heads]

Per candidate: descriptor + packed weights built against the plan's own activation buffers (after one full step, so they
hold real data), 2 warm-up launches, then `iters` launches timed with CUDA events on the launching stream."""
import argparse
import ctypes
import json
import os
import sys

import numpy as np
import torch

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=8)
    ap.add_argument("--hw", default="1080x1920")
    ap.add_argument("--keep", type=int, default=12, help="candidates per layer (best by the cycle model first)")
    ap.add_argument("--slack", type=float, default=2.0, help="only candidates modelled within this factor of the best")
    ap.add_argument("--iters", type=int, default=5)
    ap.add_argument("--layers", default=None)
    ap.add_argument("--out", default=None)
    args = ap.parse_args()
    h, w = (int(v) for v in args.hw.split("x"))
    from bench import make_net
    from lecturemath_b200 import _lib, synth
    from lecturemath_b200 import fcn_lecturenet as F
    torch.cuda.set_device(0)
    lib = _lib.lib()
    net = make_net().cuda()
    net.plan_overrides = {"no_tuned": True}                       # start from the model's choices
    frames = np.stack(list(synth.whiteboard_frames(args.batch, h, w, seed=1234)))
    plan = net.binarize_frames(frames)
    torch.cuda.synchronize()
    st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
    names = [n for n in plan.specs if args.layers is None or n in args.layers.split(",")]
    result, report = {}, []
    for name in names:
        cands = plan.conv_candidates(name)
        cands = [c for c in cands if c[0] <= cands[0][0] * args.slack][:args.keep]
        chosen = tuple(plan.specs[name]["cfg"])
        if chosen[3] is not None and not any(tuple(c[1:]) == chosen for c in cands):
            cands.append((0.0,) + chosen)
        rows = []
        for clk, S, Sy, NT, MT in cands:
            try:
                d, keep = plan.conv_variant(name, (S, Sy, NT, MT))
            except Exception as e:                                  # infeasible packing (asserts in pack_weights etc.)
                rows.append((float("inf"), S, Sy, NT, MT, "build: %s" % type(e).__name__))
                continue
            hdl = lib.am_conv_plan_create(ctypes.byref(d))
            if not hdl:
                rows.append((float("inf"), S, Sy, NT, MT, "plan_create failed"))
                continue
            for _ in range(2):
                lib.am_conv_plan_launch(hdl, st)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(args.iters):
                lib.am_conv_plan_launch(hdl, st)
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / args.iters
            lib.am_conv_plan_destroy(hdl)
            del keep
            note = "model %.3f ms" % (clk / 1.85e6)
            pool_dst = plan.specs[name].get("pool_dst")
            if pool_dst is not None and not d.pool_out:            # this packing cannot pool in its epilogue: charge the separate pass
                src = plan.specs[name]["dst"]                      # (k_maxpool2 runs at the HBM roofline: read the output, write a quarter)
                extra = args.batch * (src.H * src.Wp * src.C + pool_dst.H * pool_dst.Wp * pool_dst.C) * 2 / 6.5e12 * 1e3
                ms += extra
                note += ", + %.3f ms separate max-pool pass" % extra
            rows.append((ms, S, Sy, NT, MT, note))
        rows.sort(key=lambda r: r[0])
        base = [r for r in rows if tuple(r[1:5]) == chosen]
        best = rows[0]
        result[name] = [int(best[1]), int(best[2]), int(best[3]), int(best[4])]
        report.append((name, chosen, base[0][0] if base else None, best))
        print("%-20s model pick %s %s ms  -> best %s %.4f ms" % (name, chosen, ("%.4f" % base[0][0]) if base else "?", tuple(best[1:5]), best[0]), flush=True)
        for r in rows[:6]:
            print("      S=%d Sy=%d NT=%d MT=%d  %.4f ms  (%s)" % (r[1], r[2], r[3], r[4], r[0], r[5]), flush=True)
    tot_a = sum(r[2] for r in report if r[2]); tot_b = sum(r[3][0] for r in report)
    print("sum over layers: model picks %.3f ms -> tuned %.3f ms" % (tot_a, tot_b))
    path = args.out or F.TUNED_PATH
    table = {}
    if os.path.exists(path):
        with open(path) as f:
            table = json.load(f)
    table.setdefault(plan.arch, {}).setdefault("B%d_%dx%d" % (args.batch, h, w), {}).update(result)
    with open(path, "w") as f:
        json.dump(table, f, indent=1, sort_keys=True)
    print("wrote", path)


if __name__ == "__main__":
    main()
