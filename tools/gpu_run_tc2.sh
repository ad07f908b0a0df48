#!/bin/bash
TAG=${1:-tc2}
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "fused_matches or batch_fused or smaller_arrays or ragged or full_size_clip_config1" > gpurun_out/tests_$TAG.log 2>&1
tail -5 gpurun_out/tests_$TAG.log
timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err
tail -3 gpurun_out/bench_$TAG.err
python -c "
import json
try:
    j=json.load(open('gpurun_out/bench_$TAG.json')); print('tc  :', j['value'], j['e2e']['value'], j['roofline']['frac'])
except Exception as e: print('bench failed', e)
"
MICLOC_B200_LIB=$PWD/tools/libmicloc_b200_rt.so ROLE_NAMES=front0,front1,front2,front3,bandpass,rzcc,neuron,gram FIR_ROLES=1 timeout 300 python tools/role_timing.py 1776 2>&1 | grep -E "^rep|phase" > gpurun_out/roles_$TAG.log; cat gpurun_out/roles_$TAG.log
