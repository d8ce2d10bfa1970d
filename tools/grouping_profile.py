"""cProfile of the stage-03 drop-in on the dense 32-frame workload of tools/grouping_bench.py (where the host time goes)."""
import cProfile
import io
import os
import pstats
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    import torch
    from lecturemath_b200 import synth
    from lecturemath_b200.cc_stability_estimator import CCStabilityEstimator
    from tools.grouping_bench import stage03
    h, w, n = 1080, 1920, int(sys.argv[1]) if len(sys.argv) > 1 else 32
    masks = np.stack(list(synth.glyph_masks(n, h, w, seed=0)))
    for rep in range(2):
        est = CCStabilityEstimator(w, h, 0.85, 0.85, 85)
        est.add_frames(masks)
        torch.cuda.synchronize()
        est.device_ms = {}
        pr = cProfile.Profile()
        t = {}
        pr.enable()
        stage03(est, t)
        pr.disable()
    print(t)
    s = io.StringIO()
    pstats.Stats(pr, stream=s).sort_stats("cumulative").print_stats(45)
    print(s.getvalue())
    s = io.StringIO()
    pstats.Stats(pr, stream=s).sort_stats("tottime").print_stats(30)
    print(s.getvalue())


if __name__ == "__main__":
    main()
