#!/bin/bash
python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "segmented or wide_array or staged_taps" 2>&1 | tail -3
bash tools/gpu_run_c5.sh
python tools/c5_probe.py 4
