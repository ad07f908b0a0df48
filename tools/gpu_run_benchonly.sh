TAG=${1:-b}
( time python bench.py > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err ) 2>&1 | tail -3
python - <<PY
import json
d=json.load(open('gpurun_out/bench_$TAG.json'))
print(round(d['value']), 'clips/s frac', round(d['roofline']['frac'],4), 'tensor', round(d['roofline']['tensor_view']['frac'],3), 'e2e', round(d['e2e']['value']), 'f32', round(d['e2e']['f32']['value']), d['clocks'], d.get('rzcc_refined_clips'))
c=d.get('configs',{})
if 'error' in c: print(c['error']); print(c['trace'])
for k,v in c.items():
    if isinstance(v,dict):
        print(k, {a:(round(b,4) if isinstance(b,float) else b) for a,b in v.items() if a not in ('workload','roofline','parity','cpu_port','note')})
PY
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref_$TAG.json 2>/dev/null; cut -c1-200 gpurun_out/bench_ref_$TAG.json
