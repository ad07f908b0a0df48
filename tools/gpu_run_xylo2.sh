#!/bin/bash
TAG=${1:-x2}
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_xylo.py -q -m gpu -x --durations=5 > gpurun_out/tests_xylo_$TAG.log 2>&1; tail -15 gpurun_out/tests_xylo_$TAG.log
timeout 600 python tools/xylo_bench.py 1776 > gpurun_out/xylo_bench_$TAG.log 2>&1; cat gpurun_out/xylo_bench_$TAG.log
