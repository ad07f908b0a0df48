# tests + bench + ncu launch list + one full ncu capture of the fused kernel
mkdir -p gpurun_out
TAG=${1:-d}
nvidia-smi --query-gpu=name,clocks.max.sm,memory.total --format=csv > gpurun_out/gpu_$TAG.txt
python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/tests_$TAG.log
python bench.py --steps 5 --warmup 3 > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err
tail -3 gpurun_out/bench_$TAG.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_$TAG.csv python bench.py --steps 2 --warmup 1 --clips-per-band 888 --no-cpu > gpurun_out/bench_ncu1_$TAG.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_fused -s 1 -c 1 -o gpurun_out/prof_fused_$TAG -f python bench.py --steps 1 --warmup 1 --clips-per-band 888 --no-cpu > gpurun_out/bench_ncu2_$TAG.log 2>&1
cat gpurun_out/tests_$TAG.log
cat gpurun_out/bench_$TAG.json
ls -la gpurun_out
