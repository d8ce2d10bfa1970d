run() { AM_B200_LIB=$PWD/lecturemath_b200/libaccessmath_b200$1.so AM_B200_MATCH=$2 AM_B200_MATCH_CTAS_PER_SM=$3 timeout 300 python bench.py --no-cpu-baseline --no-cc-stage 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('lib[$1] $2 ctas/sm=$3', round(d['value'],1), round(d['ms_per_step'],3), 'nonconv', round(d['ms_per_step']-d['roofline']['conv_ms_per_step'],3))"; }
run _c2048 multi 2
run "" multi 2
run "" fused 1
run "" fused 2
run "" fused 4
run _c2048 fused 2
for g in 1 2 4; do AM_B200_MATCH_CTAS_PER_SM=$g timeout 300 python tools/cc_bench.py --batches 32 --iters 3 2>&1 | tail -1 | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('dense fused ctas/sm=$g match fps', round(d['match_frames_per_s']))"; done
AM_B200_MATCH=multi timeout 300 python tools/cc_bench.py --batches 32 --iters 3 2>&1 | tail -1 | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('dense multi match fps', round(d['match_frames_per_s']))"
