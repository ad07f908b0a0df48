for B in 3552 4736; do
python tools/xylo_bench.py $B 2>&1 | grep -E "exact"
MICLOC_B200_LIB=$PWD/tools/libmicloc_b200_vC4.so python tools/xylo_bench.py $B 2>&1 | grep -E "exact"
done
MICLOC_B200_LIB=$PWD/tools/libmicloc_b200_vC4.so python -m pytest tests/test_gpu_xylo.py -m gpu -q -x 2>&1 | tail -2
