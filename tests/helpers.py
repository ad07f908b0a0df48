"""Shared test plumbing: golden fixtures -> oracle config / engine spec, parity metrics."""
import glob
import os

import numpy as np
from scipy.signal import tf2sos

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")

SNN_CASES = sorted(os.path.basename(p)[:-4] for p in glob.glob(os.path.join(GOLDEN, "snn_*.npz")))
# one-second clips (T = 48 000) through the reference: tests/golden/make_golden_full.py
FULL_CASES = sorted(os.path.basename(p)[:-4] for p in glob.glob(os.path.join(GOLDEN, "full_c*.npz")))


def load(name):
    return dict(np.load(os.path.join(GOLDEN, name + ".npz"), allow_pickle=False))


def oracle_cfg(g):
    from oracle import oracle as O
    return O.SnnConfig(h=g["kernel"], b=g["ba_b"], a=g["ba_a"], robust_width=float(g["robust_width"]),
                       bipolar=bool(g["bipolar"]), nir=g["nir"], bf=g["bf_mat"])


def chain_spec(g, T=None):
    """Engine spec from a golden case, going through the same host code the drop-in class uses."""
    from haghighatshoarmuir2024_b200.engine import ChainSpec, neuron_alpha_params
    from scipy.signal import butter
    fs = float(g["fs"])
    T = int(g["x"].shape[0]) if T is None else T
    t = np.arange(T) / fs
    tau = float(g["tau"])
    _, a, c, L = neuron_alpha_params(t, [tau, tau])
    sos = butter(2, g["band"], btype="bandpass", output="sos", fs=fs)
    return ChainSpec(num_mic=g["x"].shape[1], stht_kernel=g["kernel"], sos=sos,
                     robust_width=int(np.ceil(float(g["robust_width"]))), bipolar=bool(g["bipolar"]),
                     neuron_decay=a, neuron_scale=c, neuron_len=L)


def rel_err(a, b):
    a = np.asarray(a); b = np.asarray(b)
    dt = np.complex128 if (np.iscomplexobj(a) or np.iscomplexobj(b)) else np.float64
    a = a.astype(dt); b = b.astype(dt)
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-300))


def spike_agreement(s, ref):
    """1 - (#positions where the rasters differ) / (#reference spikes)."""
    s = np.asarray(s).astype(np.int8); ref = np.asarray(ref).astype(np.int8)
    n_ref = int(np.count_nonzero(ref))
    return 1.0 - np.count_nonzero(s != ref) / max(n_ref, 1)


def synth_clips(g, B, T, seed, snrs_db=(-10.0, 0.0, 10.0, 20.0), int16=False):
    """Seeded noisy sine clips on the golden case's geometry (restating apply_to_template,
    micloc/snn_beamformer.py:243-275).  Returns (x [B,T,M], doa_true [B])."""
    rng = np.random.default_rng(seed)
    fs = float(g["fs"]); band = g["band"]
    t = np.arange(T) / fs
    f0 = float(np.mean(band))
    r, th = g["r_vec"], g["theta_vec"]
    xs, doas = [], []
    for i in range(B):
        doa = rng.uniform(0, 2 * np.pi)
        d = -r * np.cos(th - doa) / 340.0
        d = d - d.min()
        td = t[None, :] - d[:, None]
        td[td < 0] = 0
        x = np.interp(td.ravel(), t, np.sin(2 * np.pi * f0 * t)).reshape(td.shape).T
        snr_db = snrs_db[i % len(snrs_db)] - 10 * np.log10((fs / 2) / (band[1] - band[0]))
        x = x + np.sqrt(np.mean(x ** 2)) / np.sqrt(10 ** (snr_db / 10)) * rng.standard_normal(x.shape)
        xs.append(x); doas.append(doa)
    x = np.stack(xs)
    if int16:
        x = np.round(x / np.abs(x).max() * 16000).astype(np.int16)
    else:
        x = x.astype(np.float32)
    return x, np.asarray(doas)


# --------------------------------------------------------------------------------------
# Xylo chain
# --------------------------------------------------------------------------------------
XYLO_CASES = sorted(os.path.basename(p)[:-4] for p in glob.glob(os.path.join(GOLDEN, "xylo_*.npz")))


def xylo_network(g):
    """Quantised hidden layer of a golden Xylo case through the product's host code."""
    from haghighatshoarmuir2024_b200.xylo_snn_localization import quantize_network
    return quantize_network(list(g["bf_mats"]), g["taus"], float(g["fs"]), bool(g["bipolar"]))


def xylo_oracle_cfg(g, net=None):
    from oracle import oracle as O
    net = xylo_network(g) if net is None else net
    return O.XyloConfig(h=g["kernel"], b=g["ba_b"], a=g["ba_a"], robust_width=float(g["robust_width"]),
                        bipolar=bool(g["bipolar"]), num_mic=g["x"].shape[1], num_doa=len(g["doa_list"]),
                        w_in=net.w_in, threshold=net.threshold, dash_syn=net.dash_syn, dash_mem=net.dash_mem,
                        w_rec=net.w_rec, bias=net.bias, weight_shift_in=net.weight_shift_in,
                        weight_shift_rec=net.weight_shift_rec, max_spikes=net.max_spikes)


def xylo_engine(g, net=None, device=0):
    from scipy.signal import butter
    from haghighatshoarmuir2024_b200.xylo_snn_localization import XyloEngine
    net = xylo_network(g) if net is None else net
    fs, order = float(g["fs"]), int(g["order"])
    sos = [butter(order, band, btype="bandpass", output="sos", fs=fs) for band in g["bands"]]
    ba = [(b, a) for b, a in zip(g["ba_b"], g["ba_a"])]
    return XyloEngine(num_mic=g["x"].shape[1], stht_kernel=g["kernel"], sos_list=sos, ba_list=ba,
                      robust_width=int(np.ceil(float(g["robust_width"]))), bipolar=bool(g["bipolar"]), net=net,
                      num_doa=len(g["doa_list"]), device=device)


def xylo_synth_clips(g, B, T, seed, snrs_db=(-5.0, 5.0, 15.0), int16=False):
    """Seeded noisy chirp clips as paper_plots/target_xylo_localization.py:566-585 builds them."""
    from haghighatshoarmuir2024_b200.array_geometry import ArrayGeometry
    from haghighatshoarmuir2024_b200.xylo_snn_localization import signal_from_template
    rng = np.random.default_rng(seed)
    fs = float(g["fs"])
    geo = ArrayGeometry(g["r_vec"], g["theta_vec"])
    f_lo, f_hi = g["bands"][0]
    t = np.arange(T) / fs
    f_inst = f_lo + (f_hi - f_lo) * (t % t[-1]) / t[-1]
    src = np.sin(2 * np.pi * np.cumsum(f_inst) / fs)
    xs = []
    for i in range(B):
        sig = signal_from_template(geo, (t, src, float(rng.uniform(0, 2 * np.pi))))
        snr = 10 ** ((snrs_db[i % len(snrs_db)] - 10 * np.log10((fs / 2) / (f_hi - f_lo))) / 10)
        xs.append(sig + np.sqrt(np.mean(sig ** 2) / snr) * rng.standard_normal(sig.shape))
    x = np.stack(xs)
    if int16:
        return np.round(x / np.abs(x).max() * 12000).astype(np.int16)
    return x.astype(np.float32)
