"""oracle/time_grouping_oracle.py -- TEST INFRASTRUCTURE: wall-clock of the CPU oracle (oracle/grouping_oracle.py = the reference's
stage-03 algorithm restated) on the workload of tools/grouping_bench.py, on this box's host cores.
   python oracle/time_grouping_oracle.py [frames]"""
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import cc_oracle as CO                        # noqa: E402
from oracle.grouping_oracle import GroupingOracle         # noqa: E402
from tools.grouping_bench import stage03                  # noqa: E402
from lecturemath_b200 import synth                        # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 32
stab = CO.StabilityOracle(1920, 1080, 0.85, 0.85, 85)
for m in synth.glyph_masks(n, 1080, 1920, seed=0):
    stab.add_frame(m)
t = {}
info = stage03(GroupingOracle(stab), t)
print(json.dumps({"workload": "CPU oracle, stage 03 on %d dense 1080p glyph-mask frames (%d host cores)" % (n, os.cpu_count()), **info,
                  "cpu_oracle_ms": t, "cpu_oracle_total_ms": round(sum(t.values()), 1)}))
