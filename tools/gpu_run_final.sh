# round-end evidence on one B200: tests, default bench line, reference arm, ncu launch list, one full ncu capture
# of the fused kernel, role timing, Xylo bench, fast-FIR variant A/B
mkdir -p gpurun_out
TAG=${1:-final}
nvidia-smi --query-gpu=name,clocks.max.sm,memory.total --format=csv > gpurun_out/gpu_$TAG.txt
python -m pytest tests -m gpu -q -rs --durations=5 2>&1 | tail -20 > gpurun_out/tests_$TAG.log
cat gpurun_out/tests_$TAG.log
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
( time python bench.py > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err ) 2>&1 | tail -3
tail -2 gpurun_out/bench_$TAG.err; cat gpurun_out/bench_$TAG.json
( time python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref_$TAG.json 2> gpurun_out/bench_ref_$TAG.err ) 2>&1 | tail -3
cat gpurun_out/bench_ref_$TAG.json
MICLOC_FUSED_FIR=ffa python bench.py --steps 3 --warmup 3 --no-cpu > gpurun_out/bench_ffa_$TAG.json 2>/dev/null
python -c "
import json
for t in ('bench_$TAG','bench_ffa_$TAG'):
    d=json.load(open('gpurun_out/%s.json'%t)); print(t, round(d['value']), 'clips/s frac', round(d['roofline']['frac'],4), 'e2e', round(d['e2e']['value']))"
python bench.py --dtype i16 --steps 3 --warmup 3 --no-cpu > gpurun_out/bench_i16_$TAG.json 2>/dev/null
python -c "
import json
d=json.load(open('gpurun_out/bench_i16_$TAG.json')); print('i16 input', round(d['value']), 'clips/s e2e', round(d['e2e']['value']))"
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_$TAG.csv python bench.py --steps 2 --warmup 1 --clips-per-band 1184 --no-cpu > gpurun_out/bench_ncu1_$TAG.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_fused -s 1 -c 1 -o gpurun_out/prof_fused_$TAG -f python bench.py --steps 1 --warmup 1 --clips-per-band 1184 --no-cpu > gpurun_out/bench_ncu2_$TAG.log 2>&1
MICLOC_B200_LIB=$PWD/tools/libmicloc_b200_rt.so python tools/role_timing.py 1184 > gpurun_out/roles_$TAG.log 2>&1; tail -4 gpurun_out/roles_$TAG.log
python tools/xylo_bench.py 444 2>&1 | tee gpurun_out/xylo_bench_$TAG.log | tail -8
ls -la gpurun_out | tail -20
