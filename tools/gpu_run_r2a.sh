#!/bin/bash
# round-2 checkpoint: whole GPU suite, smoke, default bench line
TAG=${1:-r2a}
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu -rs --durations=8 > gpurun_out/tests_$TAG.log 2>&1
tail -30 gpurun_out/tests_$TAG.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
( time timeout 600 python bench.py > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err ) 2>&1 | tail -3
tail -2 gpurun_out/bench_$TAG.err; cat gpurun_out/bench_$TAG.json
