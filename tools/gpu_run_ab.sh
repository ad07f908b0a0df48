# A/B of library variants on one B200: tools/libmicloc_b200_v<X>.so through MICLOC_B200_LIB
for v in "$@"; do
  lib=$PWD/tools/libmicloc_b200_v$v.so
  [ "$v" = "main" ] && lib=$PWD/haghighatshoarmuir2024_b200/csrc/libmicloc_b200.so
  MICLOC_B200_LIB=$lib python bench.py --steps 5 --warmup 3 --no-cpu --no-extras 2>/dev/null > gpurun_out/bench_ab_$v.json
  python - <<PY
import json
d=json.load(open('gpurun_out/bench_ab_$v.json')); print('$v', round(d['value']), 'clips/s', round(d['roofline']['frac'],4), 'e2e', round(d['e2e']['value']), d['e2e'].get('matches_device_path'))
PY
done
