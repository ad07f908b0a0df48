mkdir -p gpurun_out
TAG=${1:-q}
python -m pytest tests -m gpu -x -q -rs --durations=5 2>&1 | tail -22 > gpurun_out/tests_$TAG.log
cat gpurun_out/tests_$TAG.log
