# parity tests + short bench + role timing
mkdir -p gpurun_out
TAG=${1:-x}
timeout 600 python -m pytest tests -m gpu -x -q -rs 2>&1 | tail -15 > gpurun_out/tests_$TAG.log
cat gpurun_out/tests_$TAG.log
run() { timeout 300 python bench.py --steps 3 --warmup 2 --clips-per-band 1776 --no-cpu 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('$1', round(d['value']), 'clips/s', round(d['ms_per_step'],2),'ms/step', round(d['roofline']['frac'],4), d['parity'] if 'parity' in d else '')"; }
( run "full"
MICLOC_FUSED_SKIP=0xF0 run "fir+loader only"
MICLOC_FUSED_GROUPS=1 run "groups=1" ) > gpurun_out/ab_$TAG.log 2>&1
cat gpurun_out/ab_$TAG.log
MICLOC_B200_LIB=$PWD/tools/libmicloc_b200_rt.so timeout 300 python tools/role_timing.py 1184 > gpurun_out/roles_$TAG.log 2>&1
cat gpurun_out/roles_$TAG.log
