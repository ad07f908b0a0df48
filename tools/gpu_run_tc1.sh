#!/bin/bash
# first runs of the tensor-core fused kernel: parity tests, then a short bench (both kernels)
TAG=${1:-tc1}
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "fused_matches or batch_fused or smaller_arrays or ragged or full_size_clip_config1" > gpurun_out/tests_$TAG.log 2>&1
tail -15 gpurun_out/tests_$TAG.log
timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err
tail -3 gpurun_out/bench_$TAG.err
python -c "
import json
try:
    j=json.load(open('gpurun_out/bench_$TAG.json')); print('tc  :', j['value'], j['e2e']['value'], j['roofline']['frac'])
except Exception as e: print('bench failed', e)
"
MICLOC_FUSED_FIR=ffma timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu > gpurun_out/bench_ffma_$TAG.json 2>/dev/null
python -c "
import json
try:
    j=json.load(open('gpurun_out/bench_ffma_$TAG.json')); print('ffma:', j['value'], j['e2e']['value'], j['roofline']['frac'])
except Exception as e: print('bench failed', e)
"
