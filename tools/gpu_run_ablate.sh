# marginal cost of every warp role of the default fused kernel: throughput with one role idling at the barriers
# (results of those runs are garbage; timing only)
mkdir -p gpurun_out
run() { python bench.py --steps 3 --warmup 2 --clips-per-band 1776 --no-cpu 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('$1', round(d['value']), 'clips/s', round(d['ms_per_step'],2),'ms/step')"; }
( run "all roles"
MICLOC_FUSED_SKIP=0x10 run "without band-pass"
MICLOC_FUSED_SKIP=0x20 run "without RZCC"
MICLOC_FUSED_SKIP=0x40 run "without neuron"
MICLOC_FUSED_SKIP=0x80 run "without Gram"
MICLOC_FUSED_SKIP=0xF0 run "FIR warps only"
MICLOC_FUSED_SKIP=0x0F run "serial roles only" ) 2>&1 | tee gpurun_out/ablate.log
