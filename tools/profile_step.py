"""Short, deterministic command for ncu: N steps of the fused hot path (FCN binarize + CC label/stats/crops + temporal
match) over one batch of synthetic 1080p frames resident in HBM.  No timing is reported from here (a number taken
under a profiler is never a bench value); it prints the launch count per step so `-s/-c` can be chosen.

    python tools/profile_step.py [--steps 2] [--batch 8] [--hw 1080x1920] [--cc-only]
"""
import argparse
import os
import sys

import numpy as np
import torch

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=2)
    ap.add_argument("--batch", type=int, default=8)
    ap.add_argument("--hw", default="1080x1920")
    ap.add_argument("--cc-only", action="store_true", help="dense-glyph masks through the CC stage only (BASELINE configs[3])")
    args = ap.parse_args()
    h, w = (int(v) for v in args.hw.split("x"))
    from lecturemath_b200 import synth
    torch.cuda.set_device(0)
    dev = torch.device("cuda:0")
    if args.cc_only:
        from lecturemath_b200.cc_engine import CCEngine, Estimator
        masks = np.stack(list(synth.glyph_masks(args.batch, h, w, seed=0)))
        eng = CCEngine(w, h, args.batch, device=dev)
        est = Estimator(w, h, 0.85, 0.85, 85, device=dev)
        bits = eng.pack(torch.from_numpy(masks).to(dev))
        for _ in range(args.steps):
            eng.label(bits, want_labels=False, sync=False)
            est.add_frames(eng, 0, args.batch)
        torch.cuda.synchronize()
        print("cc-only: counts", eng.read_counts().tolist(), est.state())
        return
    from bench import make_net
    from lecturemath_b200.pipeline import ContentExtractor
    net = make_net()
    ex = ContentExtractor(net, w, h, 0.85, 0.85, 85, batch=args.batch, device=dev)
    frames = torch.from_numpy(np.stack(list(synth.whiteboard_frames(args.batch, h, w, seed=1234)))).to(dev)
    for _ in range(args.steps):
        ex.launches = 0
        ex.step_device(frames, match=True)
    torch.cuda.synchronize()
    print("launches per step:", ex.launches, "state:", ex.est.state())


if __name__ == "__main__":
    main()
