# A/B of the two-group CTA (one CTA of 16 warps per SM) against two 8-warp CTAs per SM
mkdir -p gpurun_out
TAG=${1:-t}
python -m pytest tests -m gpu -x -q -rs 2>&1 | tail -6 > gpurun_out/tests_$TAG.log
cat gpurun_out/tests_$TAG.log
run() { python bench.py --steps 3 --warmup 2 --clips-per-band 1776 --no-cpu 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('$1', round(d['value']), 'clips/s', round(d['ms_per_step'],2),'ms/step', round(d['roofline']['frac'],4))"; }
( MICLOC_FUSED_GROUPS=1 run "groups=1"
run "groups=2"
MICLOC_FUSED_SKIP=0xE0 run "groups=2 fir+bandpass only"
MICLOC_FUSED_FIRBLOCKS=3 run "groups=2 firblocks=3"
MICLOC_FUSED_SKIP=0xF0 run "groups=2 fir only" ) > gpurun_out/ab_$TAG.log 2>&1
cat gpurun_out/ab_$TAG.log
MICLOC_B200_LIB=$PWD/tools/libmicloc_b200_rt.so python tools/role_timing.py 1184 > gpurun_out/roles_$TAG.log 2>&1
cat gpurun_out/roles_$TAG.log
