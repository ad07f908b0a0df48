run() { timeout 300 python bench.py --steps 3 --warmup 2 --clips-per-band $2 --no-cpu 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('$1', round(d['value']), 'clips/s', round(d['ms_per_step'],2),'ms/step', round(d['roofline']['frac'],4), d['e2e'].get('matches_device_path'))"; }
run "w descending (product)" 1776; MICLOC_FUSED_SKIP=0xF0 run "w descending FIR only" 1776
MICLOC_B200_LIB=$PWD/tools/libmicloc_b200_wasc.so run "w ascending" 1776; MICLOC_B200_LIB=$PWD/tools/libmicloc_b200_wasc.so MICLOC_FUSED_SKIP=0xF0 run "w ascending FIR only" 1776
python -m pytest tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -3
