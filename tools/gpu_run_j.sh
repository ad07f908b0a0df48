mkdir -p gpurun_out
TAG=${1:-j}
python -m pytest tests -m gpu -x -q -rs 2>&1 | tail -15 > gpurun_out/tests_$TAG.log
cat gpurun_out/tests_$TAG.log
python bench.py --steps 4 --warmup 3 --clips-per-band 3552 --no-cpu > gpurun_out/bench_${TAG}.json 2> gpurun_out/bench_${TAG}.err; tail -3 gpurun_out/bench_${TAG}.err
python -c "
import json
d=json.load(open('gpurun_out/bench_${TAG}.json')); r=d['roofline']
print(round(d['value']), 'clips/s', round(d['ms_per_step'],2),'ms/step frac', round(r['frac'],4), 'e2e', round(d['e2e']['value']), d['e2e'].get('matches_device_path'), d['clocks'])
"
