run() { timeout 300 python bench.py --steps 3 --warmup 2 --clips-per-band $2 --no-cpu 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('$1', round(d['value']), 'clips/s', round(d['ms_per_step'],2),'ms/step', round(d['roofline']['frac'],4), d['e2e'].get('matches_device_path'))"; }
run "default B=1776" 1776; MICLOC_FUSED_SKIP=0xF0 run "FIR only" 1776
