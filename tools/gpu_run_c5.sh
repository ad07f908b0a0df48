#!/bin/bash
TAG=${1:-c5}
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_parity.py -q -m gpu -x -s -k "time_segmented" --durations=5 > gpurun_out/tests_c5_$TAG.log 2>&1; tail -25 gpurun_out/tests_c5_$TAG.log
