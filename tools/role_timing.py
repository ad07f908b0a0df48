#!/usr/bin/env python
"""Busy cycles per warp role of the fused kernel (needs a library built with
`make -C haghighatshoarmuir2024_b200/csrc -B EXTRA=-DMICLOC_ROLE_TIMING`).  Runs on the GPU box."""
import ctypes
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench as Bn
from haghighatshoarmuir2024_b200 import _native as N
from haghighatshoarmuir2024_b200.montecarlo import BandSetup, SnrSweep

d, bands = Bn.load_workload()
setups = [BandSetup(band=bands[0], tau=float(d["tau_0"]), bf_mat=d["bf_0"])]
sweep = SnrSweep(setups, d["r_vec"], d["theta_vec"], Bn.FS, float(d["kernel_duration"]), Bn.T_CLIP, device=0)
B = int(sys.argv[1]) if len(sys.argv) > 1 else 888
audio, _, _ = sweep.synthesize(0, B, seed=1, snr_db_grid=Bn.SNR_GRID)
eng = sweep.engines[0]
lib = N.lib()
out = (ctypes.c_uint64 * 32)()
for rep in range(3):
    lib.micloc_snn_debug_counters(eng._h, out)
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    eng.run(audio, want_spikes=True, want_power=False)
    ev1.record()
    torch.cuda.synchronize()
    ms = ev0.elapsed_time(ev1)
    lib.micloc_snn_debug_counters(eng._h, out)
    v = list(out)
    names = os.environ.get("ROLE_NAMES", "fir0a,fir0b,fir1a,fir1b,bandpass,rzcc,neuron,gram").split(",")
    iters = (Bn.T_CLIP // 64) + 8
    print(f"rep {rep}: {ms:.3f} ms, {B / ms:.1f} clips/ms; " + ", ".join(
        f"{n}: {v[i] / max(v[8 + i], 1) / iters:.0f}" for i, n in enumerate(names)) + " busy cyc/tile")
    print(f"   CTA 0: {v[16]} clock64 cycles in {v[17]} ns -> {v[16] / max(v[17], 1) * 1e3:.0f} MHz")
    if any(v[18:]):
        print("   phase cycles per step and reporting warp:", ", ".join(f"{x / max(v[8], 1) / iters:.0f}" for x in v[18:]))

n = min(296, (B + 1) // 2)  # CTAs of the launch (148 SMs x 2)
buf = (ctypes.c_uint64 * (16 * n))()
lib.micloc_snn_debug_cta_times(eng._h, buf, n)
a = np.array(list(buf), dtype=np.float64).reshape(n, 16)
raw = np.array(list(buf), dtype=np.uint64).reshape(n, 16)
t0 = a[:, 0].min()
start, end, smid = (a[:, 0] - t0) / 1e6, (a[:, 1] - t0) / 1e6, a[:, 2].astype(int)
print("CTA start ms: max %.3f; end ms: min %.3f median %.3f max %.3f; CTAs per SM: %s" % (
    start.max(), end.min(), np.median(end), end.max(), np.bincount(np.bincount(smid))))
# FIR warps per (SM, sub-partition)
fir = {}
for i in range(n):
    for w in range(8):
        byte = (int(raw[i, 3]) >> (8 * w)) & 0xff
        role, smsp = byte & 7, (byte >> 3) & 3
        if role < int(os.environ.get("FIR_ROLES", "4")):
            fir[(smid[i], smsp)] = fir.get((smid[i], smsp), 0) + 1
print("FIR warps per (SM, sub-partition) histogram:", np.bincount(list(fir.values())), "of", 4 * len(set(smid)), "sub-partitions")
iters = Bn.T_CLIP // 64 + 8
print("mean busy cyc/tile per role over CTAs:", ", ".join(f"{nm} {a[:, 4 + i].mean() / iters:.0f}" for i, nm in enumerate(names)))
