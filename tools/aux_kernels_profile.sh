#!/bin/bash
# Runs on the GPU box: ncu launch list (time + DRAM bytes) of the kernels outside the headline step -- the > 2.5 MP path
# (k_lanczos_resize, k_bits_resize_nearest), the device PNG writer (k_png1_encode) and stage 03 (k_group_*, k_paint_items).
# usage: tools/aux_kernels_profile.sh <tag>
tag=${1:-x}
out=gpurun_out
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum
python bench.py --frame-size 3840x2160 --steps 3 --warmup 3 --no-cpu-baseline --no-cc-stage > $out/aux_plain4k_$tag.log 2>&1 && \
ncu --metrics $M --clock-control none -k regex:"k_lanczos|k_bits_resize" -s 6 -c 6 --csv --log-file $out/aux_resize_$tag.csv \
    python bench.py --frame-size 3840x2160 --steps 3 --warmup 3 --no-cpu-baseline --no-cc-stage > $out/aux_ncu4k_$tag.log 2>&1
python tools/worker_bench.py --frames 8 > $out/aux_plainw_$tag.log 2>&1 && \
ncu --metrics $M --clock-control none -k regex:k_png1 -s 8 -c 8 --csv --log-file $out/aux_png_$tag.csv python tools/worker_bench.py --frames 8 > $out/aux_ncuw_$tag.log 2>&1
python tools/grouping_bench.py --frames 32 --no-oracle > $out/aux_plaing_$tag.log 2>&1 && \
ncu --metrics $M --clock-control none -k regex:"k_group|k_paint" -c 40 --csv --log-file $out/aux_group_$tag.csv python tools/grouping_bench.py --frames 32 --no-oracle > $out/aux_ncug_$tag.log 2>&1
ls -la $out/aux_*_$tag.csv
