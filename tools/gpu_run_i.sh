mkdir -p gpurun_out
TAG=${1:-i}
python bench.py --steps 4 --warmup 3 > gpurun_out/bench_${TAG}_a.json 2> gpurun_out/bench_${TAG}_a.err; tail -2 gpurun_out/bench_${TAG}_a.err
python bench.py --steps 4 --warmup 3 --no-band-streams --no-cpu > gpurun_out/bench_${TAG}_b.json 2> gpurun_out/bench_${TAG}_b.err
python bench.py --steps 4 --warmup 3 --clips-per-band 1776 --no-cpu > gpurun_out/bench_${TAG}_c.json 2> gpurun_out/bench_${TAG}_c.err
python bench.py --steps 4 --warmup 3 --clips-per-band 14208 --no-cpu > gpurun_out/bench_${TAG}_d.json 2> gpurun_out/bench_${TAG}_d.err
for f in a b c d; do python -c "
import json,sys
d=json.load(open('gpurun_out/bench_${TAG}_$f.json')); r=d['roofline']
print('$f', round(d['value']), 'clips/s', round(d['ms_per_step'],2),'ms/step frac', round(r['frac'],4), 'e2e', round(d['e2e']['value']), d['e2e'].get('matches_device_path'), d['clocks']['sm_mhz'], d['clocks']['reasons'])
"; done
