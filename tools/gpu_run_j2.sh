mkdir -p gpurun_out
TAG=${1:-j2}
timeout 600 python -m pytest tests -m gpu -x -q -rs 2>&1 | tail -5 > gpurun_out/tests_$TAG.log
cat gpurun_out/tests_$TAG.log
run() { timeout 300 python bench.py --steps 3 --warmup 2 --clips-per-band $2 --no-cpu 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('$1', round(d['value']), 'clips/s', round(d['ms_per_step'],2),'ms/step', round(d['roofline']['frac'],4))"; }
( run "B=1776" 1776; run "B=7104" 7104 ) > gpurun_out/ab_$TAG.log 2>&1
cat gpurun_out/ab_$TAG.log
MICLOC_B200_LIB=$PWD/tools/libmicloc_b200_rt.so timeout 300 python tools/role_timing.py 1184 2>&1 | tail -1
