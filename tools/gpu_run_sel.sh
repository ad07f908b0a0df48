#!/bin/bash
# selected GPU tests: tools/gpu_run_sel.sh <tag> <pytest args...>
TAG=$1; shift
mkdir -p gpurun_out
timeout 1200 python -m pytest -q -m gpu -rs --durations=5 "$@" > gpurun_out/tests_$TAG.log 2>&1
tail -40 gpurun_out/tests_$TAG.log
