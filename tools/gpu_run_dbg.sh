#!/bin/bash
# product build: sweep batch sizes / lengths with a watchdog (a lost barrier arrival traps)
python - <<'PY' > gpurun_out/sweep_dbg.log 2>&1
import os, sys, subprocess
for B, T in [(298,1200),(300,2400),(592,1200),(700,1200),(333,700),(1000,640),(149,3000),(2000,1200)]:
    r = subprocess.run([sys.executable, "tools/wait_debug.py", str(B), str(T)], capture_output=True, text=True, timeout=200)
    print(B, T, "rc", r.returncode, (r.stdout.strip().splitlines() or ["-"])[-1][:60], (r.stderr.strip().splitlines() or ["-"])[-1][:100])
PY
cat gpurun_out/sweep_dbg.log
timeout 300 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "run_host or ragged" 2>&1 | tail -3
