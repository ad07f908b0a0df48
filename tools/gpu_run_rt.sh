#!/bin/bash
TAG=${1:-rt}
MICLOC_B200_LIB=$PWD/tools/libmicloc_b200_rt.so ROLE_NAMES=front0,front1,front2,front3,bandpass,rzcc,neuron,gram FIR_ROLES=1 timeout 300 python tools/role_timing.py 1776 2>&1 | grep -E "^rep|phase" > gpurun_out/roles_$TAG.log; cat gpurun_out/roles_$TAG.log
