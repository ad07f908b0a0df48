python tools/probe_c1.py 2>&1 | tail -30
