#!/bin/bash
# Runs on the GPU box (gpurun): launch list + ncu captures of the hot-path step; only small artefacts land in gpurun_out/.
# usage: tools/gpu_profile.sh <tag>
tag=${1:-x}
out=gpurun_out
mkdir -p $out /tmp/prof
python tools/profile_step.py > $out/plain_$tag.log 2>&1 || { echo "plain run failed"; tail -5 $out/plain_$tag.log; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $out/launches_$tag.csv python tools/profile_step.py > $out/ncu1_$tag.log 2>&1
# every conv launch of the second step, full metric set (no source import: keeps the report small)
ncu --set full --clock-control none -k regex:k_conv_gemm -s 20 -c 20 -o /tmp/prof/conv_$tag python tools/profile_step.py > $out/ncu2_$tag.log 2>&1
ncu -i /tmp/prof/conv_$tag.ncu-rep --page raw --csv > $out/conv_raw_$tag.csv 2>> $out/ncu2_$tag.log
# source-level capture of two representative layers (conv_pixels_1 = 18th conv launch, conv_up_block_1 = 16th)
ncu --set full --clock-control none --import-source on -k regex:k_conv_gemm -s 35 -c 3 -o /tmp/prof/convsrc_$tag python tools/profile_step.py > $out/ncu3_$tag.log 2>&1
ncu -i /tmp/prof/convsrc_$tag.ncu-rep --page source --csv > $out/conv_source_$tag.csv 2>> $out/ncu3_$tag.log
for f in /tmp/prof/convsrc_$tag.ncu-rep /tmp/prof/conv_$tag.ncu-rep; do
  [ -f $f ] && [ $(stat -c %s $f) -lt 12000000 ] && cp $f $out/
done
du -sh $out
# CC-only stage on dense glyph masks (BASELINE configs[3]): per-kernel time + DRAM bytes of label/stats/crops (148-frame batch, the
# third = warm iteration is summarised) and of the temporal matching (32 frames)
python tools/cc_bench.py --batches 148 --iters 1 --no-match > $out/plain_cc_$tag.log 2>&1 && \
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 40 --csv --log-file $out/launches_cc_$tag.csv python tools/cc_bench.py --batches 148 --iters 1 --no-match > $out/ncu4_$tag.log 2>&1
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 120 --csv --log-file $out/launches_match_$tag.csv python tools/cc_bench.py --batches 32 --iters 1 > $out/ncu5_$tag.log 2>&1
du -sh $out
