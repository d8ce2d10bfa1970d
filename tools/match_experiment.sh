#!/bin/bash
# Runs on the GPU box: the temporal-matching launch modes on one box, headline workload (sparse frames) and dense glyph masks.
#   AM_B200_MATCH=multi            six launches per frame (k_match_pairs ... k_match_copy)
#   AM_B200_MATCH=fused (default)  one cooperative k_match_fused launch per batch; AM_B200_MATCH_CTAS_PER_SM = 1 | 2 | 4 (default 2)
run() { AM_B200_MATCH=$1 AM_B200_MATCH_CTAS_PER_SM=$2 timeout 300 python bench.py --no-cpu-baseline --no-cc-stage 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('$1 ctas/sm=$2', round(d['value'],1), 'fps', round(d['ms_per_step'],3), 'ms/step, non-conv', round(d['ms_per_step']-d['roofline']['conv_ms_per_step'],3), 'ms')"; }
run multi 2
run fused 1
run fused 2
run fused 4
for g in 1 2 4; do AM_B200_MATCH_CTAS_PER_SM=$g timeout 300 python tools/cc_bench.py --batches 32 --iters 3 2>&1 | tail -1 | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('dense fused ctas/sm=$g match fps', round(d['match_frames_per_s']))"; done
AM_B200_MATCH=multi timeout 300 python tools/cc_bench.py --batches 32 --iters 3 2>&1 | tail -1 | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('dense multi match fps', round(d['match_frames_per_s']))"
