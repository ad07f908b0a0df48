mkdir -p gpurun_out
TAG=${1:-e}
python -m pytest tests -m gpu -x -q 2>&1 | tail -25 > gpurun_out/tests_$TAG.log
cat gpurun_out/tests_$TAG.log
python tools/xylo_bench.py 296 2>&1 | tee gpurun_out/xylo_bench_$TAG.log
