#!/bin/bash
# GPU box: ABBA comparison of programmatic dependent launch on the END-TO-END number (the device-resident pass records CUDA events between
# the conv launches, which removes the programmatic edges: only e2e can show an effect)
out=gpurun_out; mkdir -p $out
B="python bench.py --steps 40 --warmup 5 --no-cpu-baseline --no-cc-stage --no-dropin --no-gpu-reference"
i=0
for v in 0 1 1 0 0 1 1 0; do i=$((i+1))
  AM_B200_PDL=$v timeout 300 $B > $out/r02_zi_pdl${v}_$i.json 2> $out/r02_zi_pdl${v}_$i.err
  python - $out/r02_zi_pdl${v}_$i.json $v <<'P'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print("pdl", sys.argv[2], "value %.1f e2e %.1f ms %.3f conv_ms %.3f" % (d["value"], d["e2e"]["value"], d["ms_per_step"], d["roofline"]["conv_ms_per_step"]))
except Exception as e: print(sys.argv[1], "FAILED", e)
P
done
