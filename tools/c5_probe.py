"""One 10 s 64-microphone clip (BASELINE config 5) through the staged time-segmented kernels: per-kernel times come
from `ncu --metrics gpu__time_duration.sum` around this script (tools/gpu_run_c5.sh)."""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import helpers as H
from haghighatshoarmuir2024_b200.engine import SnnEngine
from haghighatshoarmuir2024_b200.montecarlo import synthesize_clips
B = int(sys.argv[1]) if len(sys.argv) > 1 else 1
g = H.load("snn_c5_linear64")
T = int(sys.argv[2]) if len(sys.argv) > 2 else 480_000
x = synthesize_clips(g["r_vec"], g["theta_vec"], 48000, T, np.full(B, 1.0), snr_lin=np.full(B, 10.0),
                     sine_freq=float(np.mean(g["band"])), mode=0, seed=3, device=0)
eng = SnnEngine(H.chain_spec(g, T), g["bf_mat"], device=0)
for _ in range(2):
    out = eng.run(x, want_spikes=True, want_power=True, fused=False)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); out = eng.run(x, want_spikes=True, want_power=True, fused=False); e1.record(); torch.cuda.synchronize()
print(f"{B} clip(s) of {T} samples: {e0.elapsed_time(e1):.2f} ms, doa {out['doa'].tolist()[:8]}")
