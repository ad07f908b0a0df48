#!/bin/bash
# full GPU suite, synccheck over the small cases, then the config-5 per-kernel launch list + timings
TAG=${1:-c5c}
bash tools/gpu_run_tests.sh $TAG
bash tools/gpu_run_sanitize.sh $TAG synccheck
bash tools/gpu_run_c5.sh
python tools/c5_probe.py 4
MICLOC_GRAM_FP32=1 MICLOC_POWER_NARROW=1 python tools/c5_probe.py 4
