"""Times conv_down_block_1 under its candidate packings in isolation (same harness as tools/autotune_fcn.py): fused-pool variants
(kPOOL2: S = Sy = 2 through shared memory; kPOOLX: S = 4, Sy = 1 through unit pairs) against the same packings without the pool.

    python tools/conv1_variants.py [--iters 10] [--layer conv_down_block_1]"""
import argparse
import ctypes
import os
import sys

import numpy as np
import torch

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--iters", type=int, default=10)
    ap.add_argument("--layer", default="conv_down_block_1")
    ap.add_argument("--once", action="store_true", help="one launch per variant (ncu)")
    args = ap.parse_args()
    from bench import make_net
    from lecturemath_b200 import _lib, synth
    torch.cuda.set_device(0)
    lib = _lib.lib()
    net = make_net().cuda()
    frames = np.stack(list(synth.whiteboard_frames(8, 1080, 1920, seed=1234)))
    plan = net.binarize_frames(frames)
    torch.cuda.synchronize()
    st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
    c = plan.specs[args.layer]["w"].shape[0]
    variants = [("S4 Sy1 unfused", (4, 1, 4 * c, 1), {"no_fused_poolx": True}), ("S4 Sy1 kPOOLX", (4, 1, 4 * c, 1), {}),
                ("S2 Sy1 unfused", (2, 1, 2 * c, 1), {"no_fused_poolx": True}), ("S2 Sy1 kPOOLX", (2, 1, 2 * c, 1), {}),
                ("S2 Sy2 unfused", (2, 2, 4 * c, 1), {"no_fused_pool2": True}), ("S2 Sy2 kPOOL2", (2, 2, 4 * c, 1), {})]
    ov0 = dict(plan.ov)
    for label, cfg, ov in variants:
        plan.ov = dict(ov0, **ov)
        d, keep = plan.conv_variant(args.layer, cfg)
        hdl = lib.am_conv_plan_create(ctypes.byref(d))
        if not hdl:
            print(label, "plan_create failed"); continue
        n_it = 1 if args.once else args.iters
        if not args.once:
            for _ in range(2):
                lib.am_conv_plan_launch(hdl, st)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(n_it):
            lib.am_conv_plan_launch(hdl, st)
        e1.record()
        torch.cuda.synchronize()
        print("%-16s pool_out=%d  %.4f ms" % (label, bool(d.pool_out), e0.elapsed_time(e1) / n_it), flush=True)
        lib.am_conv_plan_destroy(hdl)
    plan.ov = ov0


if __name__ == "__main__":
    main()
