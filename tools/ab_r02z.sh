#!/bin/bash
# GPU box: tests of the changed conv kernels, then A/B of programmatic dependent launch and of the unit-pair fused pool (kPOOLX).
out=gpurun_out; mkdir -p $out
timeout 600 python -m pytest tests/test_fcn_gpu.py tests/test_pipeline_gpu.py -x -q > $out/r02_z_tests.log 2>&1; echo "tests rc=$?"; tail -3 $out/r02_z_tests.log
B="python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-cc-stage --no-dropin --no-gpu-reference"
for rep in 1 2; do
  AM_B200_PDL=0 timeout 300 $B --layer-table $out/layers_z_pdl0_$rep.json > $out/r02_z_pdl0_$rep.json 2> $out/r02_z_pdl0_$rep.err
  AM_B200_PDL=1 timeout 300 $B --layer-table $out/layers_z_pdl1_$rep.json > $out/r02_z_pdl1_$rep.json 2> $out/r02_z_pdl1_$rep.err
  AM_B200_PDL=1 AM_B200_TUNED=$PWD/tools/tuned/poolx.json timeout 300 $B --layer-table $out/layers_z_poolx_$rep.json > $out/r02_z_poolx_$rep.json 2> $out/r02_z_poolx_$rep.err
done
for f in $out/r02_z_pdl0_1 $out/r02_z_pdl1_1 $out/r02_z_poolx_1 $out/r02_z_pdl0_2 $out/r02_z_pdl1_2 $out/r02_z_poolx_2; do
  python - $f.json <<'P'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[1], "value %.1f e2e %.1f ms %.3f conv_ms %.3f mhz %s" % (d["value"], d["e2e"]["value"], d["ms_per_step"], d["roofline"]["conv_ms_per_step"], d["clocks"]["sm_mhz"]))
except Exception as e:
    print(sys.argv[1], "FAILED", e)
P
done
