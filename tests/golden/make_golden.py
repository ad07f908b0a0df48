"""Generate the golden fixtures by running the UNMODIFIED reference (micloc, imported
from /root/reference with an empty matplotlib stub) on seeded inputs.

Run here (the container that has /root/reference); the GPU box only reads the .npz
files.  Usage:  python tests/golden/make_golden.py

Every array stored is an output of reference code:
  SNNBeamformer.design_from_template / apply_to_signal   micloc/snn_beamformer.py
  ZeroCrossingSpikeEncoder.evolve                        micloc/spike_encoder.py
  Beamformer.apply_to_signal                             micloc/beamformer.py
  find_peak_location                                     micloc/utils.py
Stage taps (q, z, spikes, vmem) are recomputed with the same scipy calls
apply_to_signal makes (snn_beamformer.py:325-364) and cross-checked against its
return value before being written.  Large float arrays keep every DEC-th row.
"""
import contextlib
import io
import os
import sys
import types

for m in ("matplotlib", "matplotlib.pyplot", "matplotlib.gridspec"):
    sys.modules[m] = types.ModuleType(m)
sys.modules["matplotlib"].gridspec = sys.modules["matplotlib.gridspec"]
sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]
sys.modules["matplotlib"].use = lambda *a, **k: None           # multiple_targets_snn.py calls use_latex() at import
sys.modules["matplotlib"].rcParams = {}
sys.modules["matplotlib.pyplot"].rc = lambda *a, **k: None
sys.path.insert(0, "/root/reference")

import numpy as np
from scipy.signal import butter, lfilter

from micloc.array_geometry import CenterCircularArray, LinearArray, Random2DArray
from micloc.beamformer import Beamformer
from micloc.snn_beamformer import SNNBeamformer
from micloc.spike_encoder import ZeroCrossingSpikeEncoder
from micloc.utils import find_peak_location

HERE = os.path.dirname(os.path.abspath(__file__))
DEC = 8
FS = 48_000


def quiet(fn, *a, **k):
    with contextlib.redirect_stdout(io.StringIO()), contextlib.redirect_stderr(io.StringIO()):
        return fn(*a, **k)


def snn_case(name, geometry, band, bipolar, G, T, source, snr_db, seed, kernel_duration=10e-3, int16=False,
             multi_targets=None, dec=DEC):
    np.random.seed(seed)
    f_mid = float(np.mean(band))
    tau = 1 / (2 * np.pi * f_mid)
    beamf = SNNBeamformer(geometry, kernel_duration, band, np.array([tau, tau]), bipolar_spikes=bipolar, fs=FS)
    t = np.arange(T) / FS
    # design template: linear chirp over the band (paper_plots/target_snn_localization.py:351-356)
    f_inst = band[0] + (band[1] - band[0]) * (t % t[-1]) / t[-1]
    chirp = np.sin(2 * np.pi * np.cumsum(f_inst) / FS)
    doa_list = np.linspace(-np.pi, np.pi, G)
    bf_mat = quiet(beamf.design_from_template, (t, chirp), doa_list)
    if source == "sine":
        src = np.sin(2 * np.pi * f_mid * t)
    elif source == "noise":
        b, a = butter(2, band, btype="bandpass", fs=FS)
        src = lfilter(b, a, np.random.randn(T))
    else:
        src = chirp
    doa = float(np.random.rand() * 2 * np.pi)
    if multi_targets is not None:
        # several simultaneous sources through the reference's own signal_multiple_targets
        # (paper_plots/multiple_targets_snn.py:87-159), speech-shaped source: white noise tilted by -6 dB/oct, band-passed
        sys.path.insert(0, "/root/reference/paper_plots")
        from multiple_targets_snn import signal_multiple_targets
        tilt = lfilter([1.0], [1.0, -0.95], np.random.randn(T))
        b_, a_ = butter(2, band, btype="bandpass", fs=FS)
        src = lfilter(b_, a_, tilt)
        doas, powers = multi_targets
        x = signal_multiple_targets(geometry, t, src, np.tile(np.asarray(doas, float), (T, 1)),
                                    np.tile(np.asarray(powers, float), (T, 1)))
        doa = float(doas[0])
    else:
        # array signal exactly as apply_to_template builds it (snn_beamformer.py:243-275)
        delays = np.asarray([geometry.delays(theta=doa, normalized=False) for _ in t]).T
        delays = delays - delays.min()
        td = t.reshape(1, -1) - delays
        td[td < t.min()] = t.min()
        x = np.interp(td.ravel(), t, src).reshape(td.shape).T
    snr = 10 ** (snr_db / 10)
    x = x + np.sqrt(np.mean(x ** 2)) / np.sqrt(snr) * np.random.randn(*x.shape)
    if int16:
        x = np.round(x / np.abs(x).max() * 12000).astype(np.int16)
        x_store = x
        x = x.astype(np.float64)
    else:
        x_store = x.astype(np.float32)
        x = x_store.astype(np.float64)          # both sides see float32-exact samples
    # reference result
    y = beamf.apply_to_signal(bf_mat, (t, x))
    power = np.mean(np.abs(y) ** 2, axis=0)
    # stage taps with the same calls
    K = beamf.kernel_length
    q = lfilter(beamf.kernel, [1], x, axis=0)
    b, a = beamf.bandpass_filter
    zh = lfilter(b, a, np.roll(x, K // 2, axis=0) + 1j * q, axis=0)
    z = np.hstack([zh.real, zh.imag])
    spikes = beamf.spk_encoder.evolve(z)
    tn = t - t[0]
    nir = (tn / tau) * np.exp(-tn / tau)
    nir = nir / np.sum(nir)
    nir = nir[: np.sum(np.cumsum(nir) < 0.999)]
    vmem = lfilter(nir, [1], spikes, axis=0)
    assert np.array_equal(vmem @ bf_mat, y), "stage taps diverge from apply_to_signal"
    rows = np.arange(0, T, dec)
    np.savez_compressed(
        os.path.join(HERE, name + ".npz"),
        kind="snn", fs=FS, band=np.asarray(band, float), bipolar=bipolar, tau=tau, kernel_duration=kernel_duration,
        r_vec=geometry.r_vec, theta_vec=geometry.theta_vec, doa_list=doa_list, doa_true=doa,
        template=chirp, x=x_store, bf_mat=bf_mat, kernel=beamf.kernel, ba_b=b, ba_a=a,
        robust_width=beamf.spk_encoder.robust_width, nir=nir,
        rows=rows, q_rows=q[rows], z_rows=z[rows], vmem_rows=vmem[rows], y_rows=y[rows],
        spikes=spikes.astype(np.int8), power=power, doa=int(np.argmax(power)),
    )
    print(name, "T", T, "G", G, "L", len(nir), "w", beamf.spk_encoder.robust_width, "spikes", int(np.abs(spikes).sum()),
          "doa", int(np.argmax(power)))


def rzcc_case():
    rng = np.random.default_rng(77)
    out = {}
    sigs = {
        "white": rng.standard_normal((2000, 3)),
        "drift": np.cumsum(rng.standard_normal((2000, 2)), axis=0) * 0.05 + 0.3,   # long monotone chains
        # flat tops in the cumsum (runs of exact zeros) with distinct peak heights.  Equal
        # heights are NOT pinned: scipy orders them with an unstable argsort, so the
        # reference's own output is unspecified there (our rule: later position wins).
        "plateau": rng.standard_normal((2000, 2)) * np.repeat(rng.random((500, 2)) > 0.4, 4, axis=0),
        "zeros": np.zeros((64, 2)),
    }
    for sname, sig in sigs.items():
        out["sig_" + sname] = sig
        for w in (1, 3, 12, 25):
            for bip in (False, True):
                enc = ZeroCrossingSpikeEncoder(fs=FS, robust_width=w, bipolar=bip)
                out[f"spk_{sname}_w{w}_b{int(bip)}"] = enc.evolve(sig).astype(np.int8)
    np.savez_compressed(os.path.join(HERE, "rzcc.npz"), **out)
    print("rzcc", len(out))


def beamformer_case():
    np.random.seed(5)
    geometry = CenterCircularArray(radius=4.5e-2, num_mic=7)
    band = [1600, 2000]
    bf = Beamformer(geometry, 10e-3, band, fs=FS)
    T, G = 2400, 24
    t = np.arange(T) / FS
    src = np.sin(2 * np.pi * 1800 * t)
    doa_list = np.linspace(-np.pi, np.pi, G)
    bf_mat, _ = quiet(bf.design_from_template, (t, src), doa_list)
    x = (np.random.randn(T, 7) * 0.3 + src[:, None]).astype(np.float32)
    y = bf.apply_to_signal(bf_mat, x.astype(np.float64))
    power = np.mean(np.abs(y) ** 2, axis=0)
    rows = np.arange(0, T, DEC)
    np.savez_compressed(os.path.join(HERE, "beamformer.npz"), fs=FS, band=np.asarray(band, float),
                        r_vec=geometry.r_vec, theta_vec=geometry.theta_vec, doa_list=doa_list, template=src,
                        x=x, bf_mat=bf_mat, rows=rows, y_rows=y[rows], power=power, doa=int(np.argmax(power)))
    print("beamformer", y.shape)


def utils_case():
    rng = np.random.default_rng(3)
    sigs = rng.random((16, 449))
    wins = [2 * ((449 // 32) // 2) + 1, 3, 5, 21]
    idx = np.array([[find_peak_location(s, w) for w in wins] for s in sigs])
    np.savez_compressed(os.path.join(HERE, "utils.npz"), sigs=sigs, wins=np.asarray(wins), idx=idx)
    print("utils", idx.shape)


def extra_cases():
    """BASELINE configs 4 and 5 at fixture size."""
    circ = CenterCircularArray(radius=4.5e-2, num_mic=7)
    # config 4: two simultaneous speech-shaped sources at +-pi/3 (multiple_targets_snn.py:307-308), 360-angle grid
    snn_case("snn_c4_multi", circ, [1500, 2500], True, 360, 4800, "noise", 15.0, 17,
             multi_targets=([np.pi / 3, -np.pi / 3], [1.0, 1.0]), dec=32)
    # config 5: 64-microphone linear and random arrays (array_resolution_linear_snn.py:126-136,
    # array_resolution_random_snn.py:114-121 with np.random.seed(1)), 512-angle grid
    lin64 = LinearArray(spacing=2 * 4.5e-2 / 64, num_mic=64, radius=4.5e-2)
    snn_case("snn_c5_linear64", lin64, [1600, 2000], True, 512, 1600, "chirp", 10.0, 18, dec=64)
    np.random.seed(1)
    rnd64 = Random2DArray(radius=4.5e-2, num_mic=64)
    snn_case("snn_c5_random64", rnd64, [1600, 2000], False, 512, 1600, "noise", 10.0, 19, dec=64)


if __name__ == "__main__":
    if "--new-only" in sys.argv:
        extra_cases()
        sys.exit(0)
    circ = CenterCircularArray(radius=4.5e-2, num_mic=7)
    snn_case("snn_c1_bipolar", circ, [1600, 2000], True, 64, 4800, "noise", 20.0, 11)
    snn_case("snn_c1_unipolar", circ, [1600, 2000], False, 64, 4800, "noise", 20.0, 12)
    snn_case("snn_band2_sine", circ, [2000, 2300], True, 57, 4800, "sine", 0.0, 13)
    snn_case("snn_band3_i16", circ, [2300, 2600], True, 33, 2400, "sine", -8.0, 14, int16=True)
    lin = LinearArray(spacing=2 * 4.5e-2 / 16, num_mic=16, radius=4.5e-2)
    snn_case("snn_linear16", lin, [1600, 2000], True, 40, 2400, "chirp", 10.0, 15)
    snn_case("snn_k20ms", circ, [2300, 2600], False, 24, 2400, "noise", 10.0, 16, kernel_duration=20e-3)
    extra_cases()
    if "--new-only" not in sys.argv:
        rzcc_case()
        beamformer_case()
        utils_case()
