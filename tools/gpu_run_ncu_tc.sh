#!/bin/bash
TAG=${1:-tc}
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_fused_tc -s 1 -c 1 -o gpurun_out/prof_$TAG -f python bench.py --steps 1 --warmup 1 --clips-per-band 1184 --no-cpu > gpurun_out/bench_ncu_$TAG.log 2>&1
tail -3 gpurun_out/bench_ncu_$TAG.log
ls -la gpurun_out/prof_$TAG.ncu-rep
