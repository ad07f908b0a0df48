mkdir -p gpurun_out
TAG=${1:-l2}
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -rs -k "fast_fir" 2>&1 | tail -15 > gpurun_out/tests_$TAG.log
cat gpurun_out/tests_$TAG.log
run() { timeout 300 python bench.py --steps 3 --warmup 2 --clips-per-band $2 --no-cpu 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('$1', round(d['value']), 'clips/s', round(d['ms_per_step'],2),'ms/step', round(d['roofline']['frac'],4), d['e2e'].get('matches_device_path'))"; }
( MICLOC_FUSED_FIR=ffa run "fast FIR thirds B=1776" 1776; run "default B=1776" 1776; MICLOC_FUSED_FIR=ffa MICLOC_FUSED_SKIP=0xF0 run "fast FIR, FIR only" 1776 ) > gpurun_out/ab_$TAG.log 2>&1
cat gpurun_out/ab_$TAG.log
MICLOC_FUSED_FIR=ffa ROLE_NAMES=fir0,fir1,fir2,fir3,bandpass,rzcc,neuron,gram MICLOC_B200_LIB=$PWD/tools/libmicloc_b200_rt.so timeout 300 python tools/role_timing.py 1184 2>&1 | tail -1
