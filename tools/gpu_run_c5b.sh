python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "config5 or segmented or stream or golden" 2>&1 | tail -4
python -m pytest tests/test_gpu_stream.py -m gpu -q -x 2>&1 | tail -2
bash tools/gpu_run_c5.sh
python tools/c5_probe.py 4
