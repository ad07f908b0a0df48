# tests + default bench line + role timing + role ablation of the fused kernel (one B200)
mkdir -p gpurun_out
TAG=${1:-s}
python -m pytest tests -m gpu -x -q -rs --durations=5 2>&1 | tail -22 > gpurun_out/tests_$TAG.log
cat gpurun_out/tests_$TAG.log
( time python bench.py > gpurun_out/bench_${TAG}.json 2> gpurun_out/bench_${TAG}.err ) 2>&1 | tail -4
tail -3 gpurun_out/bench_${TAG}.err
cat gpurun_out/bench_${TAG}.json
MICLOC_B200_LIB=$PWD/tools/libmicloc_b200_rt.so python tools/role_timing.py 888 > gpurun_out/roles_$TAG.log 2>&1
cat gpurun_out/roles_$TAG.log
bash tools/gpu_run_n.sh > gpurun_out/ablate_$TAG.log 2>&1
cat gpurun_out/ablate_$TAG.log
