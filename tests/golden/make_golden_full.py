"""Full-length golden fixtures (T = 48 000, one second at 48 kHz) from the UNMODIFIED reference, imported from
/root/reference with an empty matplotlib stub.  Run here:  python tests/golden/make_golden_full.py

  full_c1.npz            BASELINE configs[0]: noisy wideband source, band 1600-2000 Hz, G = 64, bipolar;
                         bf_mat from the reference's design_from_template (chirp template, 1 s)
  full_c2_b{0,1,2}.npz   one clip per band of configs[1] (sine source, G = 449, the matrices of bench_c2_bf.npz,
                         themselves the reference's design_from_template output)
  full_multiband.npz     the live demo's frame (micloc/localization_demo_snn.py:125-193): int32 `T x 8` wav frame,
                         last channel dropped, order-1 Butterworth filterbank, one SNNBeamformer per band (sine
                         template at the band centre), power summed over the bands, argmax; + the periodic-ML
                         estimators of micloc/xylo_snn_localization.py:400-444 applied to that power pattern

Inputs are int16 / int32 integers so that the reference (float64) and the device (float32) see the same samples.
Every stored output is what reference code returned: apply_to_signal's y (decimated rows), the power, the argmax,
and the spike raster / membrane rows recomputed with the very calls apply_to_signal makes and cross-checked against
its return value.  These pin what the T <= 4800 fixtures cannot: the neuron kernel's normalisation over T
(snn_beamformer.py:342-356) and spike agreement over a whole second.
"""
import contextlib
import io
import os
import sys
import types

for m in ("matplotlib", "matplotlib.pyplot", "matplotlib.gridspec"):
    sys.modules[m] = types.ModuleType(m)
sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]
sys.path.insert(0, "/root/reference")

import numpy as np
from scipy.signal import butter, lfilter

from micloc.array_geometry import CenterCircularArray
from micloc.filterbank import ButterworthFilterbank
from micloc.snn_beamformer import SNNBeamformer

HERE = os.path.dirname(os.path.abspath(__file__))
FS = 48_000
DEC = 64


def quiet(fn, *a, **k):
    with contextlib.redirect_stdout(io.StringIO()), contextlib.redirect_stderr(io.StringIO()):
        return fn(*a, **k)


def array_signal(geometry, t, src, doa):
    """apply_to_template's array signal (snn_beamformer.py:243-268), noise added by the caller."""
    delays = np.asarray([geometry.delays(theta=doa, normalized=False)] * len(t)).T
    delays = delays - delays.min()
    td = t.reshape(1, -1) - delays
    td[td < t.min()] = t.min()
    return np.interp(td.ravel(), t, src).reshape(td.shape).T


def chain_taps(beamf, bf_mat, t, x):
    """Reference result + the stage taps recomputed with apply_to_signal's own calls (snn_beamformer.py:325-368)."""
    y = beamf.apply_to_signal(bf_mat, (t, x))
    K = beamf.kernel_length
    q = lfilter(beamf.kernel, [1], x, axis=0)
    b, a = beamf.bandpass_filter
    zh = lfilter(b, a, np.roll(x, K // 2, axis=0) + 1j * q, axis=0)
    spikes = beamf.spk_encoder.evolve(np.hstack([zh.real, zh.imag]))
    tau = beamf.tau_vec[0]
    tn = t - t[0]
    nir = (tn / tau) * np.exp(-tn / tau)
    nir = nir / np.sum(nir)
    nir = nir[: np.sum(np.cumsum(nir) < 0.999)]
    vmem = lfilter(nir, [1], spikes, axis=0)
    assert np.array_equal(vmem @ bf_mat, y), "stage taps diverge from apply_to_signal"
    return y, q, spikes, vmem, nir


def save_case(name, geometry, beamf, bf_mat, doa_list, band, tau, t, x_int, extra=None):
    x = x_int.astype(np.float64)
    y, q, spikes, vmem, nir = chain_taps(beamf, bf_mat, t, x)
    power = np.mean(np.abs(y) ** 2, axis=0)
    rows = np.arange(0, len(t), DEC)
    yrows = np.arange(0, len(t), DEC * (8 if bf_mat.shape[1] > 64 else 1))
    b, a = beamf.bandpass_filter
    np.savez_compressed(
        os.path.join(HERE, name + ".npz"),
        kind="snn", fs=FS, band=np.asarray(band, float), bipolar=beamf.bipolar_spikes, tau=tau,
        kernel_duration=beamf.kernel_duration, r_vec=geometry.r_vec, theta_vec=geometry.theta_vec, doa_list=doa_list,
        x=x_int, bf_mat=bf_mat, kernel=beamf.kernel, ba_b=b, ba_a=a, robust_width=beamf.spk_encoder.robust_width,
        nir=nir, rows=rows, q_rows=q[rows], vmem_rows=vmem[rows], yrows=yrows, y_rows=y[yrows],
        spikes=spikes.astype(np.int8),
        power=power, doa=int(np.argmax(power)), **(extra or {}))
    print(name, "T", len(t), "G", bf_mat.shape[1], "L", len(nir), "spikes", int(np.abs(spikes).sum()), "doa",
          int(np.argmax(power)), flush=True)
    return power


def quantise(x, full_scale, dtype):
    return np.round(x / np.abs(x).max() * full_scale).astype(dtype)


def c1_case():
    np.random.seed(101)
    geometry = CenterCircularArray(radius=4.5e-2, num_mic=7)
    band = [1600, 2000]
    tau = 1 / (2 * np.pi * np.mean(band))
    beamf = SNNBeamformer(geometry, 10e-3, band, np.array([tau, tau]), bipolar_spikes=True, fs=FS)
    t = np.arange(FS) / FS
    f_inst = band[0] + (band[1] - band[0]) * (t % t[-1]) / t[-1]
    chirp = np.sin(2 * np.pi * np.cumsum(f_inst) / FS)
    doa_list = np.linspace(-np.pi, np.pi, 64)
    bf_mat = quiet(beamf.design_from_template, (t, chirp), doa_list)
    b, a = butter(2, band, btype="bandpass", fs=FS)
    src = lfilter(b, a, np.random.randn(len(t)))                 # tests/test_snn_hilbert_localization.py:59-64
    doa = float(np.random.rand() * 2 * np.pi)
    x = array_signal(geometry, t, src, doa)
    x = x + np.sqrt(np.mean(x ** 2)) / np.sqrt(10 ** (10.0 / 10)) * np.random.randn(*x.shape)
    save_case("full_c1", geometry, beamf, bf_mat, doa_list, band, tau, t, quantise(x, 12000, np.int16),
              extra=dict(doa_true=doa))


def c2_cases():
    d = np.load(os.path.join(HERE, "bench_c2_bf.npz"))
    geometry = CenterCircularArray(radius=4.5e-2, num_mic=7)
    t = np.arange(FS) / FS
    for i, band in enumerate(d["bands"]):
        np.random.seed(200 + i)
        tau = float(d[f"tau_{i}"])
        beamf = SNNBeamformer(geometry, 10e-3, list(band), np.array([tau, tau]), bipolar_spikes=True, fs=FS)
        src = np.sin(2 * np.pi * float(np.mean(band)) * t)       # target_snn_localization.py:439-441
        doa = float(np.random.rand() * 2 * np.pi)
        x = array_signal(geometry, t, src, doa)
        snr_db = (0.0, 8.0, 20.0)[i] - 10 * np.log10((FS / 2) / (band[1] - band[0]))     # :382,449
        x = x + np.sqrt(np.mean(x ** 2)) / np.sqrt(10 ** (snr_db / 10)) * np.random.randn(*x.shape)
        save_case(f"full_c2_b{i}", geometry, beamf, d[f"bf_{i}"], d["doa_list"], band, tau, t,
                  quantise(x, 12000, np.int16), extra=dict(doa_true=doa))


def multiband_case():
    """localization_demo_snn.py:52-98 (set-up) and :125-193 (one frame)."""
    np.random.seed(303)
    geometry = CenterCircularArray(radius=4.5e-2, num_mic=7)
    bands = [[1600, 2000], [2000, 2300], [2300, 2600]]
    # (the reference's trimmed_periodic_ml indexes with arange(-G/4, G/4 + 1) - argmax and raises IndexError when the
    #  argmax exceeds 3G/4: the source direction is chosen inside)
    T = 12_000                                                   # recording_duration 0.25 s
    t = np.arange(T) / FS
    doa_list = np.linspace(-np.pi, np.pi, 64)
    beamfs, bf_mats, taus = [], [], []
    for band in bands:
        f_mid = np.mean(band)
        tau = 1 / (2 * np.pi * f_mid)
        beamf = SNNBeamformer(geometry, 10e-3, band, [tau, tau], bipolar_spikes=True, fs=FS)
        bf_mats.append(quiet(beamf.design_from_template, (t, np.sin(2 * np.pi * f_mid * t)), doa_list))
        beamfs.append(beamf); taus.append(tau)
    fb = ButterworthFilterbank(freq_bands=bands, order=1, fs=FS)
    # a wideband source over all three bands + noise, as an int32 wav frame with an 8th all-zero channel
    b, a = butter(2, [1500, 2700], btype="bandpass", fs=FS)
    src = lfilter(b, a, np.random.randn(T))
    doa = 0.9
    x = array_signal(geometry, t, src, doa)
    x = x + np.sqrt(np.mean(x ** 2)) / np.sqrt(10 ** (5.0 / 10)) * np.random.randn(*x.shape)
    frame = np.concatenate([quantise(x, 2 ** 28, np.int32), np.zeros((T, 1), np.int32)], axis=1)
    # --- the frame as run() processes it ---
    data = np.asarray(frame[:, :-1], dtype=np.float64)
    data_filt = fb.evolve(sig_in=data)
    power_grid = 0
    powers, spikes = [], []
    for chan, bf, beamf in zip(data_filt, bf_mats, beamfs):
        y, _, s, _, _ = chain_taps(beamf, bf, t, chan)
        p = np.mean(np.abs(y) ** 2, axis=0)
        powers.append(p); spikes.append(s.astype(np.int8))
        power_grid = power_grid + p
    doa_index = int(np.argmax(power_grid))
    # estimators of xylo_snn_localization.py:424-444 on this pattern
    periodic_ml = float(np.angle(np.mean(power_grid * np.exp(1j * doa_list))))
    num_doa = len(doa_list) // 2
    rng_idx = np.arange(-num_doa // 2, num_doa // 2 + 1) - doa_index
    trimmed = float(np.angle(np.mean(power_grid[rng_idx] * np.exp(1j * doa_list[rng_idx]))))
    np.savez_compressed(
        os.path.join(HERE, "full_multiband.npz"),
        fs=FS, bands=np.asarray(bands, float), taus=np.asarray(taus), bipolar=True, kernel_duration=10e-3,
        r_vec=geometry.r_vec, theta_vec=geometry.theta_vec, doa_list=doa_list, frame=frame,
        bf_mats=np.asarray(bf_mats), kernel=beamfs[0].kernel,
        fb_b=np.asarray([b for b, _ in fb.ba_list]), fb_a=np.asarray([a for _, a in fb.ba_list]),
        robust_widths=np.asarray([bm.spk_encoder.robust_width for bm in beamfs]),
        powers=np.asarray(powers), spikes=np.asarray(spikes), power_grid=power_grid, doa=doa_index, doa_true=doa,
        periodic_ml=periodic_ml, trimmed_periodic_ml=trimmed)
    print("full_multiband doa", doa_index, "true", doa, "periodic_ml", periodic_ml, "trimmed", trimmed, flush=True)


if __name__ == "__main__":
    which = sys.argv[1:] or ["c1", "c2", "multiband"]
    if "c1" in which: c1_case()
    if "c2" in which: c2_cases()
    if "multiband" in which: multiband_case()
