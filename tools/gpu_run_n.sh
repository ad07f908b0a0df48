mkdir -p gpurun_out
for m in 0x00 0xF0 0x0F 0xE0 0xD0 0xB0 0x70 0x0A 0x05; do
MICLOC_FUSED_SKIP=$m python bench.py --steps 3 --warmup 2 --clips-per-band 1184 --no-cpu > gpurun_out/skip_$m.json 2>/dev/null
python -c "
import json
d=json.load(open('gpurun_out/skip_$m.json'))
print('skip $m', round(d['value']), 'clips/s', round(d['ms_per_step'],2),'ms/step')
"
done
