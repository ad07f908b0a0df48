python tools/xylo_bench.py 3552 2>&1 | grep exact
python tools/xylo_bench.py 3552 2>&1 | grep exact
MICLOC_B200_LIB=$PWD/tools/libmicloc_b200_vQ.so python -m pytest tests/test_gpu_xylo.py -m gpu -q -x 2>&1 | tail -2
MICLOC_B200_LIB=$PWD/tools/libmicloc_b200_vQ.so python tools/xylo_bench.py 3552 2>&1 | grep exact
MICLOC_B200_LIB=$PWD/tools/libmicloc_b200_vQ.so python tools/xylo_bench.py 3552 2>&1 | grep exact
