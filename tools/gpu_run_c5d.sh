#!/bin/bash
python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "config5 or segmented or wide_array" 2>&1 | tail -4
bash tools/gpu_run_c5.sh
python tools/c5_probe.py 4
MICLOC_CHAIN_SEG_V1=1 python tools/c5_probe.py 4
