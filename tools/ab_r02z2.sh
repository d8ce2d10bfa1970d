#!/bin/bash
out=gpurun_out; mkdir -p $out /tmp/prof
python tools/conv1_variants.py > $out/conv1_variants.log 2>&1; cat $out/conv1_variants.log | tail -8
python tools/conv1_variants.py --layer conv_down_block_2 > $out/conv2_variants.log 2>&1; cat $out/conv2_variants.log | tail -8
ncu --set full --clock-control none --import-source on -k regex:k_conv_gemm -s 20 -c 6 -o /tmp/prof/conv1v python tools/conv1_variants.py --once > $out/ncu_conv1v.log 2>&1
ncu -i /tmp/prof/conv1v.ncu-rep --page raw --csv > $out/conv1v_raw.csv 2>> $out/ncu_conv1v.log
ncu -i /tmp/prof/conv1v.ncu-rep --page source --csv > $out/conv1v_source.csv 2>> $out/ncu_conv1v.log
ls -la /tmp/prof $out/conv1v_raw.csv $out/conv1v_source.csv
