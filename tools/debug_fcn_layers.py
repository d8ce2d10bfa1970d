"""GPU debugging aid: run an FCNPlan on the device and on the CPU emulation (tests/emulate_fcn.py) with the same
weights and report the max abs difference of every intermediate buffer, in execution order.
    python tools/debug_fcn_layers.py [tiny|full] [rowrun=1|0]"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from lecturemath_b200.fcn_lecturenet import FCNPlan   # noqa: E402
from tests.emulate_fcn import emulate_plan            # noqa: E402
from tests.test_fcn_host_logic import golden_net      # noqa: E402


def buffers(plan):
    names = [("x0", plan.x0)]
    for i in range(5):
        names += [("d%d" % (i + 1), plan.d[i]), ("p%d" % (i + 1), plan.p[i])]
    names.append(("mid", plan.mid))
    for i in range(4, -1, -1):
        names += [("t%d" % (i + 1), plan.t[i]), ("u%d" % (i + 1), plan.u[i])]
    names += [("diff", plan.diff), ("px1", plan.px1), ("px2", plan.px2)]
    return names


def main():
    tag = sys.argv[1] if len(sys.argv) > 1 else "tiny"
    rowrun = bool(int(sys.argv[2])) if len(sys.argv) > 2 else True
    z = np.load(os.path.join(os.path.dirname(__file__), "..", "tests", "golden", "fcn_forward.npz"))
    net = golden_net(tag, z)
    frame = z["frame_bgr"]
    H, W = frame.shape[:2]
    cpu = FCNPlan(net.params, 1, H, W, torch.device("cpu"), rowrun=rowrun)
    emulate_plan(cpu, frame[None])
    dev = torch.device("cuda:0")
    gpu = FCNPlan(net.params, 1, H, W, dev, rowrun=rowrun)
    gpu.frames.copy_(torch.from_numpy(frame[None]))
    gpu.run(torch.cuda.current_stream().cuda_stream, want_others=True)
    torch.cuda.synchronize()
    print("%-6s %12s %12s" % ("buffer", "max|gpu-emu|", "max|emu|"))
    for (name, a), (_, b) in zip(buffers(gpu), buffers(cpu)):
        ga, eb = a.view().float().cpu(), b.view().float()
        print("%-6s %12.5f %12.5f" % (name, (ga - eb).abs().max().item(), eb.abs().max().item()))
    for name in ("heads", "logits"):
        ga, eb = getattr(gpu, name).float().cpu(), getattr(cpu, name).float()
        print("%-6s %12.5f %12.5f" % (name, (ga - eb).abs().max().item(), eb.abs().max().item()))
    ref = z[tag + "_logit"]
    print("max |sigmoid(gpu) - sigmoid(reference)| = %.5f" % np.abs(1 / (1 + np.exp(-gpu.logits[0].cpu().numpy())) - 1 / (1 + np.exp(-ref))).max())


if __name__ == "__main__":
    main()
