import os, sys, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import helpers as H, bench as Bn
from haghighatshoarmuir2024_b200.engine import SnnEngine
from haghighatshoarmuir2024_b200.montecarlo import synthesize_clips
FS, T = 48000, 48000
t = np.arange(T) / FS
rng = np.random.default_rng(5)
g = H.load("snn_c1_bipolar")
B = 1776
f_lo, f_hi = map(float, g["band"])
chirp = np.sin(2 * np.pi * np.cumsum(f_lo + (f_hi - f_lo) * t / t[-1]) / FS)
snr = 10 ** ((Bn.SNR_GRID[np.arange(B) % 11] - 10 * np.log10((FS / 2) / (f_hi - f_lo))) / 10)
doa = rng.uniform(0, 2 * np.pi, B)
eng = SnnEngine(H.chain_spec(g, T), g["bf_mat"], device=0)
def run(name, x, **kw):
    eng.enable_timing(True)
    for _ in range(3):
        out = eng.run(x, want_spikes=True, want_power=True, refine=False, **kw)
    torch.cuda.synchronize()
    ms, n = eng.last_kernel_ms()
    print(f"{name:28s} {ms / n:8.3f} ms/launch  {B / (ms / n) * 1e3:9.0f} clips/s  density {float(out['spikes'][:32].ne(0).float().mean()):.4f} flags {int(out['flags'].sum())}  amax {float(x.abs().max()):.3g}")
run("chirp + snr grid", synthesize_clips(g["r_vec"], g["theta_vec"], FS, T, doa, snr_lin=snr, source=chirp, mode=0, seed=71))
run("chirp + snr 100", synthesize_clips(g["r_vec"], g["theta_vec"], FS, T, doa, snr_lin=np.full(B, 100.0), source=chirp, mode=0, seed=71))
run("sine 2000 + snr grid", synthesize_clips(g["r_vec"], g["theta_vec"], FS, T, doa, snr_lin=snr, sine_freq=2000.0, mode=0, seed=71))
run("sine 1800 + snr grid", synthesize_clips(g["r_vec"], g["theta_vec"], FS, T, doa, snr_lin=snr, sine_freq=1800.0, mode=0, seed=71))
run("noise only", torch.randn(B, T, 7, device="cuda"))
x = synthesize_clips(g["r_vec"], g["theta_vec"], FS, T, doa, snr_lin=snr, source=chirp, mode=0, seed=71)
run("chirp, no power", x)
os.environ["MICLOC_FUSED_FIR"] = "ffma"
run("chirp + grid, ffma kernel", x)
del os.environ["MICLOC_FUSED_FIR"]
# does the float32 band-pass output ever reach exact zero in digital silence?
y = x[:4, :24000].clone(); y[1, 3000:15000] = 0
tp = eng.run_taps(y, want=("z", "spikes"))
z = tp["z"][1].cpu().numpy()
for c in (0, 3, 7, 10):
    zz = z[3000:15000, c]
    nz = np.nonzero(zz == 0)[0]
    print("channel", c, "exact zeros:", len(nz), "first at", (nz[0] if len(nz) else None), "min |z| nonzero", np.abs(zz[zz != 0]).min() if np.any(zz != 0) else 0, "tail values", zz[-4:])
