#!/bin/bash
TAG=${1:-xf}
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_xylo_chain_f64 -s 1 -c 1 -o gpurun_out/prof_$TAG -f python tools/xylo_front_only.py 1776 > gpurun_out/ncu_$TAG.log 2>&1
tail -3 gpurun_out/ncu_$TAG.log
ls -la gpurun_out/prof_$TAG.ncu-rep
