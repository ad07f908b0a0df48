bash tools/gpu_run_tests.sh c5l | tail -4
python tools/c5_probe.py 256 48000
python tools/c5_probe.py 4
