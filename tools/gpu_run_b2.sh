mkdir -p gpurun_out
TAG=${1:-b2}
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -rs 2>&1 | tail -5 > gpurun_out/tests_$TAG.log
cat gpurun_out/tests_$TAG.log
run() { timeout 300 python bench.py --steps 3 --warmup 2 --clips-per-band 1776 --no-cpu 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('$1', round(d['value']), 'clips/s', round(d['ms_per_step'],2),'ms/step', round(d['roofline']['frac'],4))"; }
( MICLOC_FUSED_LAYOUT=2 run "layout 2 (balanced)"
MICLOC_FUSED_LAYOUT=0 run "layout 0 (mixed)"
MICLOC_FUSED_LAYOUT=1 run "layout 1 (apart)"
MICLOC_FUSED_LAYOUT=2 MICLOC_FUSED_SKIP=0xF0 run "layout 2 fir+loader only"
MICLOC_FUSED_LAYOUT=2 MICLOC_FUSED_SKIP=0x07 run "layout 2 serial+loader only"
MICLOC_FUSED_LAYOUT=1 MICLOC_FUSED_SKIP=0x07 run "layout 1 serial+loader only" ) > gpurun_out/ab_$TAG.log 2>&1
cat gpurun_out/ab_$TAG.log
MICLOC_FUSED_LAYOUT=2 MICLOC_B200_LIB=$PWD/tools/libmicloc_b200_rt.so timeout 300 python tools/role_timing.py 1184 2>&1 | tail -3 > gpurun_out/roles_$TAG.log
cat gpurun_out/roles_$TAG.log
