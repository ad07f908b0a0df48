mkdir -p gpurun_out
TAG=${1:-g}
python -m pytest tests/test_gpu_xylo.py -m gpu -x -q -rs 2>&1 | tail -25 > gpurun_out/tests_$TAG.log
cat gpurun_out/tests_$TAG.log
python tools/xylo_bench.py 444 2>&1 | tee gpurun_out/xylo_bench_$TAG.log
python tools/xylo_bench.py 888 2>&1 | tee -a gpurun_out/xylo_bench_$TAG.log
ncu --set full --clock-control none --import-source on -k regex:k_xylo_lif -s 1 -c 1 -o gpurun_out/prof_lif_$TAG -f python tools/xylo_bench.py 444 > gpurun_out/ncu_lif_$TAG.log 2>&1
