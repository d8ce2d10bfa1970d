#!/bin/bash
# GPU box: what the driver runs at round end (GPU tests, smoke, default bench) on the committed build
tag=${1:-zj}
out=gpurun_out; mkdir -p $out
timeout 900 python -m pytest tests -x -q -m gpu > $out/r02_${tag}_tests.log 2>&1; echo "tests rc=$?"; tail -2 $out/r02_${tag}_tests.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > $out/r02_${tag}_smoke.log 2>&1; echo "smoke rc=$?"; tail -1 $out/r02_${tag}_smoke.log
timeout 600 python bench.py > $out/r02_${tag}_bench.json 2> $out/r02_${tag}_bench.err; echo "bench rc=$?"
python - $out/r02_${tag}_bench.json <<'P'
import json,sys
d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
e=d["e2e_dropin"]
print("value %.1f e2e %.1f ms %.3f conv_ms %.3f frac %.3f | dropin seq %.0f s1 %.0f s2 %.0f fused %.0f | launches %d" % (d["value"], d["e2e"]["value"], d["ms_per_step"], d["roofline"]["conv_ms_per_step"], d["roofline"]["frac"], e["value"], e["stage01_frames_per_s"], e["stage02_frames_per_s"], e["fused"]["value"], d["gpu_launches"]))
P
