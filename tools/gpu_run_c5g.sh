#!/bin/bash
python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "segmented or wide_array or staged_taps" 2>&1 | tail -3
bash tools/gpu_run_c5.sh
python tools/c5_probe.py 4
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_chain_blk -s 1 -c 1 -o gpurun_out/prof_c5blk -f python tools/c5_probe.py 1 > gpurun_out/ncu_c5blk.log 2>&1
tail -2 gpurun_out/ncu_c5blk.log
