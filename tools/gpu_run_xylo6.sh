python -m pytest tests/test_gpu_xylo.py -m gpu -q -x 2>&1 | tail -2
python tools/xylo_bench.py 3552 2>&1 | grep -E "exact|LIF"
timeout 300 ncu --set full --clock-control none -k regex:k_xylo_lif_mma -s 1 -c 1 -o gpurun_out/prof_lifmma -f python tools/xylo_bench.py 1776 > /dev/null 2>&1
ls -la gpurun_out/prof_lifmma.ncu-rep
