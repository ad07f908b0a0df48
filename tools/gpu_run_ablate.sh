mkdir -p gpurun_out
run() { python bench.py --steps 3 --warmup 2 --clips-per-band 1184 --no-cpu 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('$1', round(d['value']), 'clips/s', round(d['ms_per_step'],2),'ms/step')"; }
MICLOC_FUSED_FIRBLOCKS=3 run "firblocks=3 (all roles)"
MICLOC_FUSED_FIRBLOCKS=3 MICLOC_FUSED_SKIP=0xE0 run "firblocks=3 + bandpass only"
MICLOC_FUSED_FIRBLOCKS=3 MICLOC_FUSED_SKIP=0x80 run "firblocks=3, no gram"
MICLOC_FUSED_FIRBLOCKS=3 MICLOC_FUSED_SKIP=0xC0 run "firblocks=3, no neuron/gram"
MICLOC_FUSED_FIRBLOCKS=9 run "firblocks=9 (all roles)"
run "full"
