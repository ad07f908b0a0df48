#!/bin/bash
# compute-sanitizer pass over the GPU parity suite (small cases): tools/gpu_run_sanitize.sh <tag> <tool> [pytest -k expr]
TAG=${1:-s}; TOOL=${2:-memcheck}; KEXPR=${3:-not full_size}
( time timeout ${SAN_TIMEOUT:-420} compute-sanitizer --tool $TOOL --print-limit 30 --error-exitcode 9 \
    python -m pytest tests -q -m gpu -k "$KEXPR" -p no:cacheprovider > gpurun_out/sanitize_${TOOL}_$TAG.log 2>&1 ) 2>&1 | tail -3
echo "rc=$?"
grep -c "========= " gpurun_out/sanitize_${TOOL}_$TAG.log
grep -m 40 "=========\|passed\|failed" gpurun_out/sanitize_${TOOL}_$TAG.log | cut -c1-220 | tail -40
