"""Golden fixtures of the reference's INPUT SYNTHESIS, captured from the unmodified reference code:

  mode 0  SNNBeamformer.apply_to_template (micloc/snn_beamformer.py:243-275): the array signal it hands to
          apply_to_signal is captured by replacing that one bound method on the instance; snr_db = 400 makes its
          AWGN vanish (1e-20 relative).
  mode 1  signal_multiple_targets (paper_plots/multiple_targets_snn.py:87-159), imported and called as is.

Run here (the container that has /root/reference): python tests/golden/make_golden_synth.py
"""
import os
import sys
import types

for m in ("matplotlib", "matplotlib.pyplot", "matplotlib.gridspec"):
    sys.modules[m] = types.ModuleType(m)
sys.modules["matplotlib"].gridspec = sys.modules["matplotlib.gridspec"]
sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]
sys.modules["matplotlib"].use = lambda *a, **k: None
sys.modules["matplotlib"].rcParams = {}
sys.modules["matplotlib.pyplot"].rc = lambda *a, **k: None
sys.path.insert(0, "/root/reference")
sys.path.insert(0, "/root/reference/paper_plots")

import numpy as np
from scipy.signal import butter, lfilter

from micloc.array_geometry import CenterCircularArray, LinearArray
from micloc.snn_beamformer import SNNBeamformer

HERE = os.path.dirname(os.path.abspath(__file__))
FS = 48_000


def captured_template_signal(geometry, template, snr_db=400.0):
    band = [1600, 2000]
    tau = 1 / (2 * np.pi * 1800.0)
    beamf = SNNBeamformer(geometry, 10e-3, band, np.array([tau, tau]), bipolar_spikes=True, fs=FS)
    beamf.apply_to_signal = lambda bf_mat, sig_in_vec: sig_in_vec        # capture what apply_to_template synthesised
    t, x = beamf.apply_to_template(bf_mat=None, template=template, snr_db=snr_db)
    return np.asarray(t), np.asarray(x)


def main():
    out = {}
    rng = np.random.default_rng(2024)
    T = 4801                                   # np.arange(min, max, 1/fs) drops the last sample: 4800 frames
    t = np.arange(T) / FS
    geo7 = CenterCircularArray(radius=4.5e-2, num_mic=7)
    geo16 = LinearArray(spacing=2 * 4.5e-2 / 16, num_mic=16, radius=4.5e-2)
    sine = np.sin(2 * np.pi * 2000.0 * t)
    f_inst = 1600 + 400 * (t % t[-1]) / t[-1]
    chirp = np.sin(2 * np.pi * np.cumsum(f_inst) / FS)
    b, a = butter(2, [1600, 2000], btype="bandpass", fs=FS)
    noise = lfilter(b, a, rng.standard_normal(T))
    cases = [("sine7", geo7, sine, 0.7), ("sine7b", geo7, sine, -2.1), ("chirp7", geo7, chirp, 2.9),
             ("noise7", geo7, noise, 4.0), ("chirp16", geo16, chirp, 1.1)]
    for name, geo, sig, doa in cases:
        tt, x = captured_template_signal(geo, (t, sig, doa))
        out[f"m0_{name}_x"] = x
        out[f"m0_{name}_src"] = np.interp(tt, t, sig)       # the template on the clip grid (snn_beamformer.py:246-247)
        out[f"m0_{name}_doa"] = np.float64(doa)
        out[f"m0_{name}_r"] = geo.r_vec
        out[f"m0_{name}_theta"] = geo.theta_vec
    # mode 1: two and three simultaneous targets, band-limited noise source
    from multiple_targets_snn import signal_multiple_targets
    T1 = 4800
    t1 = np.arange(T1) / FS
    src = lfilter(b, a, rng.standard_normal(T1))
    for name, doas, gains in [("two", [np.pi / 3, -np.pi / 3], [1.0, 0.7]), ("three", [0.3, 2.0, -1.2], [1.0, 0.5, 1.5])]:
        doa_ts = np.tile(np.asarray(doas), (T1, 1))
        pow_ts = np.tile(np.asarray(gains), (T1, 1))
        x = signal_multiple_targets(geo7, t1, src, doa_ts, pow_ts)
        out[f"m1_{name}_x"] = x
        out[f"m1_{name}_src"] = src
        out[f"m1_{name}_doa"] = np.asarray(doas)
        out[f"m1_{name}_gain"] = np.asarray(gains)
    out["r7"], out["theta7"] = geo7.r_vec, geo7.theta_vec
    np.savez_compressed(os.path.join(HERE, "synth.npz"), **out)
    print({k: np.shape(v) for k, v in out.items()})


if __name__ == "__main__":
    main()
