# stagger sweep of the two-group fused kernel (full and FIR-only)
mkdir -p gpurun_out
TAG=${1:-v}
python -m pytest tests/test_gpu_parity.py -m gpu -x -q -rs 2>&1 | tail -4 > gpurun_out/tests_$TAG.log
cat gpurun_out/tests_$TAG.log
run() { python bench.py --steps 3 --warmup 2 --clips-per-band 1776 --no-cpu 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('$1', round(d['value']), 'clips/s', round(d['ms_per_step'],2),'ms/step', round(d['roofline']['frac'],4))"; }
( for s in 0 800 1600 2400 3200 4000 5000 6000; do MICLOC_FUSED_STAGGER=$s run "full stagger=$s"; done
for s in 0 1000 2000 3000; do MICLOC_FUSED_SKIP=0xF0 MICLOC_FUSED_STAGGER=$s run "fir-only stagger=$s"; done ) > gpurun_out/stagger_$TAG.log 2>&1
cat gpurun_out/stagger_$TAG.log
