#!/bin/bash
# GPU box: the non-headline bench lines of the final build (sustained 10k frames, 720p, 4K, dense masks), plus the new full-size test
tag=${1:-zf}
out=gpurun_out; mkdir -p $out
timeout 300 python -m pytest tests/test_fcn_gpu.py -x -q -m gpu -k full_size > $out/r02_${tag}_tests.log 2>&1; echo "tests rc=$?"; tail -1 $out/r02_${tag}_tests.log
X="--no-cpu-baseline --no-cc-stage --no-dropin --no-gpu-reference"
timeout 300 python bench.py --steps 1250 --warmup 5 $X > $out/r02_${tag}_bench_10k_frames.json 2> $out/r02_${tag}_10k.err; echo "10k rc=$?"
timeout 300 python bench.py --frame-size 1280x720 $X > $out/r02_${tag}_bench_720p.json 2> $out/r02_${tag}_720p.err; echo "720p rc=$?"
timeout 300 python bench.py --frame-size 3840x2160 $X > $out/r02_${tag}_bench_4k.json 2> $out/r02_${tag}_4k.err; echo "4k rc=$?"
timeout 300 python bench.py --masks glyph $X > $out/r02_${tag}_bench_glyph.json 2> $out/r02_${tag}_glyph.err; echo "glyph rc=$?"
timeout 300 python tools/worker_bench.py > $out/r02_${tag}_worker_bench.json 2> $out/r02_${tag}_worker.err; echo "worker rc=$?"
for f in 10k_frames 720p 4k glyph; do python - $out/r02_${tag}_bench_$f.json <<'P'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[1].split("bench_")[-1], "value %.1f e2e %.1f ms %.3f conv_ms %.3f frac %.3f (%s) mhz %s" % (d["value"], d["e2e"]["value"], d["ms_per_step"], d["roofline"]["conv_ms_per_step"], d["roofline"]["frac"], d["roofline"]["peak_kind"][:40], d["clocks"]["sm_mhz"]))
except Exception as e: print(sys.argv[1], "FAILED", e)
P
done
cat $out/r02_${tag}_worker_bench.json | tail -1 | cut -c1-400
