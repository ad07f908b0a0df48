# quick post-change check on one B200: all GPU tests, smoke, one short bench line
python -m pytest tests -m gpu -q 2>&1 | tail -4
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -1
python bench.py --steps 3 --warmup 3 --no-cpu 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print(round(d['value']), 'clips/s', round(d['roofline']['frac'],4), 'e2e', round(d['e2e']['value']), d['e2e'].get('matches_device_path'), d['clocks'])"
