#!/bin/bash
python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "tensor_core_stht or staged_taps or wide_array or segmented" 2>&1 | tail -6
bash tools/gpu_run_c5.sh
python tools/c5_probe.py 4
