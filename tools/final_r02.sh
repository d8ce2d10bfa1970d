#!/bin/bash
# GPU box: full validation of HEAD + the profile set that goes under profiles/ (tag from $1)
tag=${1:-fin}
out=gpurun_out; mkdir -p $out
timeout 900 python -m pytest tests -x -q -m gpu > $out/r02_${tag}_tests.log 2>&1; echo "tests rc=$?"; tail -2 $out/r02_${tag}_tests.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > $out/r02_${tag}_smoke.log 2>&1; echo "smoke rc=$?"; tail -1 $out/r02_${tag}_smoke.log
timeout 600 python bench.py --layer-table $out/layers_${tag}.json > $out/r02_${tag}_bench.json 2> $out/r02_${tag}_bench.err; echo "bench rc=$?"
timeout 600 python bench.py --impl reference > $out/r02_${tag}_bench_reference.json 2> $out/r02_${tag}_bench_reference.err; echo "ref rc=$?"
bash tools/gpu_profile.sh $tag > $out/r02_${tag}_profile.log 2>&1; echo "profile rc=$?"
cut -c1-400 $out/r02_${tag}_bench.json
