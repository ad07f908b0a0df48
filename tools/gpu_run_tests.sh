#!/bin/bash
TAG=${1:-t}
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/tests_$TAG.log 2>&1
tail -25 gpurun_out/tests_$TAG.log
