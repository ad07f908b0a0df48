#!/bin/bash
MICLOC_B200_LIB=$PWD/tools/libmicloc_b200_san.so bash tools/gpu_run_sanitize.sh san synccheck
bash tools/gpu_run_benchonly.sh b6
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_chain_seg -s 1 -c 1 -o gpurun_out/prof_c5chain -f python tools/c5_probe.py 1 > gpurun_out/ncu_c5chain.log 2>&1
tail -2 gpurun_out/ncu_c5chain.log
