python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "staged_taps or wide_array or tensor_core_stht or segmented or dropin or hilbert" 2>&1 | tail -3
python tools/c5_probe.py 256 48000
MICLOC_CHAIN_SEG_V1=1 python tools/c5_probe.py 256 48000
MICLOC_CHAIN_SEG_V1=1 MICLOC_STHT_FP32=1 MICLOC_GRAM_FP32=1 MICLOC_POWER_NARROW=1 python tools/c5_probe.py 256 48000
