python -m pytest tests/test_gpu_xylo.py -m gpu -q -x 2>&1 | tail -8
python tools/xylo_bench.py 3552 2>&1 | tail -6
MICLOC_XYLO_LIF_ADDS=1 python tools/xylo_bench.py 3552 2>&1 | tail -6
