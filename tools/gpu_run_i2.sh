mkdir -p gpurun_out
TAG=${1:-i2}
run() { timeout 300 python bench.py --steps 3 --warmup 2 --clips-per-band $2 --no-cpu 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('$1', round(d['value']), 'clips/s', round(d['ms_per_step'],2),'ms/step', round(d['roofline']['frac'],4))"; }
( MICLOC_FUSED_LAYOUT=3 run "layout 3" 1776; MICLOC_FUSED_LAYOUT=2 run "layout 2" 1776; MICLOC_FUSED_LAYOUT=3 run "layout 3 B=7104" 7104; MICLOC_FUSED_LAYOUT=2 run "layout 2 B=7104" 7104 ) > gpurun_out/ab_$TAG.log 2>&1
cat gpurun_out/ab_$TAG.log
