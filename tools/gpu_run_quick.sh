# quick post-change check on one B200: fused-kernel parity tests, smoke, one short bench line
TAG=${1:-q}
python -m pytest tests/test_gpu_parity.py tests/test_gpu_xylo.py tests/test_gpu_multiband.py -m gpu -q -x 2>&1 | tail -6
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -1
python bench.py --steps 5 --warmup 3 --no-cpu --no-extras 2>/dev/null > gpurun_out/bench_$TAG.json; python - <<PY
import json
d=json.load(open('gpurun_out/bench_$TAG.json')); print(round(d['value']), 'clips/s', round(d['roofline']['frac'],4), 'e2e', round(d['e2e']['value']), d['e2e'].get('matches_device_path'), d['clocks'])
PY
