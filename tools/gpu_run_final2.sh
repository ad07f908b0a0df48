# round-2 final evidence on one B200
mkdir -p gpurun_out
TAG=${1:-fin}
python -m pytest tests -m gpu -q -rs --durations=5 2>&1 | tail -16 > gpurun_out/tests_$TAG.log; tail -4 gpurun_out/tests_$TAG.log
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -1
( time python bench.py > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err ) 2>&1 | tail -3
python - <<PY
import json
d=json.load(open('gpurun_out/bench_$TAG.json'))
print(round(d['value']), 'clips/s frac', round(d['roofline']['frac'],4), 'e2e', round(d['e2e']['value']), 'f32', round(d['e2e']['f32']['value']), d['clocks'], d.get('rzcc_refined_clips'))
c=d.get('configs',{})
if 'error' in c: print(c['error']); print(c['trace'])
for k,v in c.items():
    if isinstance(v,dict):
        print(k, {a:(round(b,4) if isinstance(b,float) else b) for a,b in v.items() if a not in ('workload','roofline','roofline_front_end','parity','cpu_port','note')})
PY
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref_$TAG.json 2>/dev/null
python bench.py --dtype i16 --steps 3 --warmup 3 --no-cpu --no-extras > gpurun_out/bench_i16_$TAG.json 2>/dev/null
MICLOC_FUSED_FIR=ffma python bench.py --steps 3 --warmup 3 --no-cpu --no-extras > gpurun_out/bench_ffma_$TAG.json 2>/dev/null
python - <<PY
import json
for t in ('bench_i16_$TAG','bench_ffma_$TAG'):
    d=json.load(open('gpurun_out/%s.json'%t)); print(t, round(d['value']), 'clips/s frac', round(d['roofline']['frac'],4), 'e2e', round(d['e2e']['value']))
PY
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_$TAG.csv python bench.py --steps 2 --warmup 1 --clips-per-band 1184 --no-cpu --no-extras > /dev/null 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_fused_tc -s 1 -c 1 -o gpurun_out/prof_tc_$TAG -f python bench.py --steps 1 --warmup 1 --clips-per-band 1184 --no-cpu --no-extras > /dev/null 2>&1
ls -la gpurun_out/prof_tc_$TAG.ncu-rep
