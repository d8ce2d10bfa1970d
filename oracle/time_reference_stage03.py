"""Container-only (needs /root/reference): wall-clock of the UNMODIFIED reference's stage 02 + stage 03 estimator methods on the
dense 1080p glyph-mask workload of tools/grouping_bench.py, for the record in DESIGN.md.   python oracle/time_reference_stage03.py 32"""
import json
import os
import sys
import time

import numpy as np

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
from oracle.gen_golden import import_reference           # noqa: E402
from tools.grouping_bench import stage03                  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 32
Labeler, CCStabilityEstimator, *_ = import_reference()
from lecturemath_b200 import synth                        # noqa: E402
masks = list(synth.glyph_masks(n, 1080, 1920, seed=0))
est = CCStabilityEstimator(1920, 1080, 0.85, 0.85, 85, False)
t0 = time.perf_counter()
for m in masks:
    est.add_frame(m, True)
t02 = time.perf_counter() - t0
t = {}
info = stage03(est, t)
print(json.dumps({"workload": "UNMODIFIED reference, %d dense 1080p glyph-mask frames, container CPU (%d cores)" % (n, os.cpu_count()),
                  "stage02_ms_per_frame": round(1000 * t02 / n, 1), **info, "stage03_ms": t, "stage03_total_ms": round(sum(t.values()), 1)}))
