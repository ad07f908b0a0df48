"""Golden fixtures of the Xylo front end, produced by the UNMODIFIED reference code.

rockpool / xylosim / samna are not installed here, so `micloc.xylo_snn_localization` cannot run
its network part.  Its front end and post-processing are plain numpy/scipy, though: this script
puts empty stand-in modules for rockpool (and matplotlib) into sys.modules so that the module
IMPORTS, then calls the reference's own, unmodified
    Demo.spike_encoding          micloc/xylo_snn_localization.py:315-356
    Demo.extract_rate            :379-398
    Demo.estimate_doa_from_rate  :400-444
    signal_from_template         :44-71
    find_peak_location           micloc/utils.py:84-121
as unbound functions on a plain namespace that carries the attributes those methods read
(beamfs, filterbank, bipolar_spikes, freq_bands, doa_list, fs) -- all of them built by the
reference's own classes.  The XyloSim hidden layer itself has NO reference output here
(parity unpinned).

Run here (needs /root/reference):  python tests/golden/make_golden_xylo.py
"""
import contextlib
import io
import os
import sys
import types

for m in ("matplotlib", "matplotlib.pyplot", "rockpool", "rockpool.nn", "rockpool.nn.modules",
          "rockpool.nn.combinators", "rockpool.devices", "rockpool.devices.xylo",
          "rockpool.devices.xylo.syns61201", "rockpool.transform", "micloc.record", "micloc.visualizer"):
    sys.modules[m] = types.ModuleType(m)
sys.modules["rockpool.nn.modules"].LinearTorch = sys.modules["rockpool.nn.modules"].LIFBitshiftTorch = None
sys.modules["rockpool.nn.modules"].LIFTorch = None
sys.modules["rockpool.nn.combinators"].Sequential = None
for name in ("config_from_specification", "mapper", "xa2_devkit_utils", "XyloSamna", "XyloSim"):
    setattr(sys.modules["rockpool.devices.xylo.syns61201"], name, None)
sys.modules["rockpool.transform"].quantize_methods = None
sys.modules["micloc.record"].AudioRecorder = None
sys.modules["micloc.visualizer"].Visualizer = None
sys.path.insert(0, "/root/reference")

import numpy as np

from micloc.array_geometry import CenterCircularArray
from micloc.filterbank import ButterworthFilterbank
from micloc.snn_beamformer import SNNBeamformer
from micloc.utils import find_peak_location
from micloc.xylo_snn_localization import Demo, signal_from_template

HERE = os.path.dirname(os.path.abspath(__file__))
FS = 48_000


def quiet(fn, *a, **k):
    with contextlib.redirect_stdout(io.StringIO()), contextlib.redirect_stderr(io.StringIO()):
        return fn(*a, **k)


def xylo_case(name, bands, order, bipolar, G, T, T_design, snr_db, seed, int16):
    """Front end of Demo for `bands` (Demo.__init__, :107-152, with `order` as in :150 /
    paper_plots/snn_localization_benchmark.py:153) on one noisy clip."""
    np.random.seed(seed)
    geometry = CenterCircularArray(radius=4.5e-2, num_mic=7)
    bands = np.asarray(bands, dtype=np.float64).reshape(-1, 2)
    doa_list = np.linspace(-np.pi, np.pi, G)
    beamfs, bf_mats, taus = [], [], []
    for fr in bands:
        f_mid = np.mean(fr)
        tau = 1 / (2 * np.pi * f_mid)
        beamf = SNNBeamformer(geometry=geometry, kernel_duration=10e-3, freq_range=fr, tau_vec=[tau, tau],
                              bipolar_spikes=bipolar, fs=FS)
        tt = np.arange(0, T_design / FS, step=1 / FS)
        bf_mats.append(quiet(beamf.design_from_template, template=(tt, np.sin(2 * np.pi * f_mid * tt)), doa_list=doa_list))
        beamfs.append(beamf)
        taus.append([tau, tau])
    fb = ButterworthFilterbank(freq_bands=bands, order=order, fs=FS)
    ns = types.SimpleNamespace(beamfs=beamfs, filterbank=fb, bipolar_spikes=bipolar, freq_bands=bands,
                               doa_list=doa_list, fs=FS)
    # test clip as in paper_plots/target_xylo_localization.py:566-585 (chirp over the first band, AWGN)
    t = np.arange(T) / FS
    f_lo, f_hi = bands[0]
    f_inst = f_lo + (f_hi - f_lo) * (t % t[-1]) / t[-1]
    src = np.sin(2 * np.pi * np.cumsum(f_inst) / FS)
    doa = float(np.random.rand() * 2 * np.pi)
    sig = signal_from_template(template=(t, src, doa), geometry=geometry)
    snr = 10 ** (snr_db / 10)
    x = sig + np.sqrt(np.mean(sig ** 2) / snr) * np.random.randn(*sig.shape)
    if int16:
        x_store = np.round(x / np.abs(x).max() * 12000).astype(np.int16)
    else:
        x_store = x.astype(np.float32)
    x = x_store.astype(np.float64)                  # both sides see exactly these samples
    spikes_in = Demo.spike_encoding(ns, x)
    # post-processing goldens on a synthetic hidden raster (the real one needs XyloSim)
    rng = np.random.default_rng(seed)
    N = len(bands) * G
    lam = 0.02 + 0.3 * np.exp(-0.5 * ((np.arange(N) % G - 0.6 * G) / (0.05 * G)) ** 2)
    raster = rng.poisson(lam, size=(600, N)).astype(np.int64)
    rate = Demo.extract_rate(ns, raster)
    est = {m: float(Demo.estimate_doa_from_rate(ns, rate, m)) for m in ("peak", "periodic_ml", "trimmed_periodic_ml")}
    win = 2 * ((G // 32) // 2) + 1
    peak_idx = find_peak_location(sig_in=rate / rate.max(), win_size=win)
    np.savez_compressed(
        os.path.join(HERE, name + ".npz"),
        kind="xylo", fs=FS, bands=bands, order=order, bipolar=bipolar, kernel_duration=10e-3,
        r_vec=geometry.r_vec, theta_vec=geometry.theta_vec, doa_list=doa_list, doa_true=doa,
        kernel=beamfs[0].kernel, robust_width=beamfs[0].spk_encoder.robust_width,
        ba_b=np.stack([b for b, a in fb.ba_list]), ba_a=np.stack([a for b, a in fb.ba_list]),
        taus=np.asarray(taus), bf_mats=np.stack(bf_mats), x=x_store, sig_clean_rows=sig[::16],
        spikes_in=spikes_in.astype(np.int8),
        post_raster=raster.astype(np.uint8), post_rate=rate, post_peak=est["peak"], post_periodic_ml=est["periodic_ml"],
        post_trimmed_periodic_ml=est["trimmed_periodic_ml"], post_win=win, post_peak_index=int(peak_idx),
    )
    print(name, "bands", len(bands), "order", order, "N_in", spikes_in.shape[1], "G", G, "T", T,
          "input spikes", int(spikes_in.sum()), "w", beamfs[0].spk_encoder.robust_width)


if __name__ == "__main__":
    # config 3 (paper_plots/target_xylo_localization.py:407-444): one band [1000, 2000], order-1 filterbank,
    # G = 449, bipolar (28 inputs) and unipolar (14 inputs)
    xylo_case("xylo_c3_bipolar", [[1000, 2000]], 1, True, 449, 4800, 2400, 10.0, 21, int16=False)
    xylo_case("xylo_c3_unipolar", [[1000, 2000]], 1, False, 449, 4800, 2400, 5.0, 22, int16=True)
    # three bands + order-2 filterbank (paper_plots/snn_localization_benchmark.py:153, 556-561)
    xylo_case("xylo_3band_o2", [[1600, 2000], [2000, 2300], [2300, 2600]], 2, True, 33, 3600, 2400, 10.0, 23, int16=True)
