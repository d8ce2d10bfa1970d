#!/bin/bash
# GPU box: FCN tests of the in-tree library, then A/B of bench.py between it and lecturemath_b200/libaccessmath_b200_old.so (AM_B200_LIB)
out=gpurun_out; mkdir -p $out
timeout 600 python -m pytest tests/test_fcn_gpu.py tests/test_pipeline_gpu.py tests/test_resize_gpu.py -x -q -m gpu > $out/r02_zh_tests.log 2>&1; echo "tests rc=$?"; tail -1 $out/r02_zh_tests.log
B="python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-cc-stage --no-dropin --no-gpu-reference"
for rep in 1 2; do
  AM_B200_LIB=$PWD/lecturemath_b200/libaccessmath_b200_old.so timeout 300 $B --layer-table $out/layers_zh_old_$rep.json > $out/r02_zh_old_$rep.json 2> $out/r02_zh_old_$rep.err
  timeout 300 $B --layer-table $out/layers_zh_new_$rep.json > $out/r02_zh_new_$rep.json 2> $out/r02_zh_new_$rep.err
done
for f in old_1 new_1 old_2 new_2; do python - $out/r02_zh_$f.json <<'P'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[1].split("zh_")[-1], "value %.1f e2e %.1f ms %.3f conv_ms %.3f" % (d["value"], d["e2e"]["value"], d["ms_per_step"], d["roofline"]["conv_ms_per_step"]))
except Exception as e: print(sys.argv[1], "FAILED", e)
P
done
python tools/compare_layers.py $out/layers_zh_old_2.json $out/layers_zh_new_2.json
