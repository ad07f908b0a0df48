mkdir -p gpurun_out
TAG=${1:-r2c}
python -m pytest tests -m gpu -q -x -k "overflow or rockpool" 2>&1 | tail -15
( time python bench.py --steps 3 --warmup 3 > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err ) 2>&1 | tail -3
tail -3 gpurun_out/bench_$TAG.err
python - <<PY
import json
d=json.load(open('gpurun_out/bench_$TAG.json'))
print(round(d['value']), 'clips/s frac', round(d['roofline']['frac'],4), 'e2e', round(d['e2e']['value']), 'f32', round(d['e2e']['f32']['value']), d['clocks'])
c=d.get('configs',{})
if 'error' in c: print(c['error']); print(c['trace'])
for k,v in c.items():
    if isinstance(v,dict):
        print(k, {a:(round(b,4) if isinstance(b,float) else b) for a,b in v.items() if a not in ('workload','roofline','roofline_front_end','parity','cpu_port')})
        print('   ', v.get('roofline') or v.get('roofline_front_end')); print('   ', v.get('parity'), v.get('cpu_port'))
PY
