"""Accuracy of the FCN binarizer against the fp32 oracle on the same GPU (cuDNN fp32, TF32 off): max probability error and
mask disagreement on synthetic whiteboard frames with the reference's seed-0 random init.  Test infrastructure (imports
oracle/); used to compare kernel variants (AM_B200_LIB=...).   python oracle/check_fcn_accuracy.py [--hw 1080x1920] [--frames 2]"""
import argparse
import json
import os
import sys

import numpy as np
import torch

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--hw", default="1080x1920")
    ap.add_argument("--frames", type=int, default=2)
    args = ap.parse_args()
    h, w = (int(v) for v in args.hw.split("x"))
    from lecturemath_b200 import synth
    from lecturemath_b200.configuration import Configuration
    from lecturemath_b200.fcn_lecturenet import FCN_LectureNet
    from oracle import fcn_oracle as FO
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.manual_seed(0)
    net = FCN_LectureNet.CreateFromConfig(Configuration.from_file(os.path.join(REPO, "tests", "golden", "fcn_full.conf")), 3, False).eval().cuda()
    frames = np.stack(list(synth.whiteboard_frames(args.frames, h, w, seed=1234)))
    plan = net.binarize_frames(frames)
    torch.cuda.synchronize()
    sd = {k: v.cuda() for k, v in net.state_dict().items()}
    out = []
    for f in range(args.frames):
        ref, _, _ = FO.forward(sd, FO.prepare_image(frames[f][:, :, ::-1]).cuda())
        p, pr = torch.sigmoid(plan.logits[f]), torch.sigmoid(ref[0, 0])
        d = (p - pr).abs()
        ink, ink_ref = (p * 255).to(torch.uint8) < 128, (pr * 255).to(torch.uint8) < 128
        out.append({"max_prob_err": d.max().item(), "mean_prob_err": d.mean().item(), "p999_prob_err": d.flatten().float().kthvalue(int(0.999 * d.numel())).values.item(),
                    "logit_rms_err": (plan.logits[f] - ref[0, 0]).pow(2).mean().sqrt().item(), "mask_disagreement": (ink != ink_ref).float().mean().item()})
    print(json.dumps({"lib": os.environ.get("AM_B200_LIB", "default"), "hw": [h, w], "frames": out}))


if __name__ == "__main__":
    main()
