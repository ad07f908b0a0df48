"""CPU tests of the Xylo chain: the oracle's front end against goldens produced by the reference's
own Demo.spike_encoding (tests/golden/make_golden_xylo.py), the host-side mirrors, and the integer
LIF restatement against an independent numpy restatement (XyloSim itself is absent: parity unpinned)."""
import types

import numpy as np
import pytest

import helpers as H
from oracle import oracle as O


@pytest.mark.parametrize("name", H.XYLO_CASES)
def test_front_end_matches_reference_spike_encoding_bit_exact(name):
    g = H.load(name)
    cfg = H.xylo_oracle_cfg(g)
    spk, sgn = O.xylo_encode(cfg, g["x"].astype(np.float64))
    assert spk.shape == g["spikes_in"].shape
    assert np.array_equal(spk, g["spikes_in"])
    CT = sgn.shape[1]
    if bool(g["bipolar"]):
        assert np.array_equal(sgn, g["spikes_in"][:, :CT] - g["spikes_in"][:, CT:])
    else:
        assert np.array_equal(sgn, g["spikes_in"])


@pytest.mark.parametrize("name", H.XYLO_CASES)
def test_signal_from_template_mirror(name):
    from haghighatshoarmuir2024_b200.array_geometry import ArrayGeometry
    from haghighatshoarmuir2024_b200.xylo_snn_localization import signal_from_template
    g = H.load(name)
    fs = float(g["fs"]); T = g["x"].shape[0]
    t = np.arange(T) / fs
    f_lo, f_hi = g["bands"][0]
    f_inst = f_lo + (f_hi - f_lo) * (t % t[-1]) / t[-1]
    src = np.sin(2 * np.pi * np.cumsum(f_inst) / fs)
    sig = signal_from_template(ArrayGeometry(g["r_vec"], g["theta_vec"]), (t, src, float(g["doa_true"])))
    np.testing.assert_array_equal(sig[::16], g["sig_clean_rows"])


@pytest.mark.parametrize("name", H.XYLO_CASES)
def test_rate_and_estimators_match_reference(name):
    from haghighatshoarmuir2024_b200.xylo_snn_localization import Demo
    g = H.load(name)
    G, F = len(g["doa_list"]), len(g["bands"])
    raster = g["post_raster"].astype(np.int64)
    ns = types.SimpleNamespace(freq_bands=g["bands"], doa_list=g["doa_list"], fs=float(g["fs"]))
    rate = Demo.extract_rate(ns, raster)
    np.testing.assert_array_equal(rate, g["post_rate"])
    for m in ("peak", "periodic_ml", "trimmed_periodic_ml"):
        assert float(Demo.estimate_doa_from_rate(ns, rate, m)) == float(g["post_" + m])
    with pytest.raises(ValueError):
        Demo.estimate_doa_from_rate(ns, rate, "median")
    # oracle: integer estimators == the reference's float ones
    counts = raster.sum(axis=0).astype(np.int32)
    r2, doa, doa_peak = O.xylo_rate_doa(counts, G, F, raster.shape[0], float(g["fs"]), int(g["post_win"]))
    np.testing.assert_allclose(r2, rate, rtol=1e-15)
    assert g["doa_list"][doa] == float(g["post_peak"])
    assert doa_peak == int(g["post_peak_index"])


@pytest.mark.parametrize("name", H.XYLO_CASES)
def test_quantised_network_facts(name):
    g = H.load(name)
    net = H.xylo_network(g)
    F, G = len(g["bands"]), len(g["doa_list"])
    M = g["x"].shape[1]
    n_in = 2 * M * F * (2 if bool(g["bipolar"]) else 1)
    assert net.w_in.shape == (n_in, F * G) and net.w_in.dtype == np.int8
    assert np.abs(net.w_in.astype(int)).max() == 127
    # tau * fs = 1/(2 pi f_mid) * fs in [3, 5.1] -> dash 2; w_rec = -0.1/N rounds to zero
    assert np.all(net.dash_syn == 2) and np.all(net.dash_mem == 2)
    assert net.w_rec is None
    assert np.all(net.threshold == int(np.round(net.scale)))
    if bool(g["bipolar"]):
        assert np.array_equal(net.w_in[: n_in // 2], -net.w_in[n_in // 2:])
    # block diagonal over bands
    blk = net.w_in[: 2 * M * F].reshape(F, 2 * M, F, G)
    for f in range(F):
        for f2 in range(F):
            if f != f2:
                assert not blk[f, :, f2, :].any()


def numpy_lif(spikes_in, w_in, thr, dash_syn, dash_mem, max_spikes=31, w_rec=None):
    """Independent vectorised restatement of the hidden-layer dynamics (see oracle/micloc_oracle.c)."""
    T, _ = spikes_in.shape
    N = w_in.shape[1]
    isyn = np.zeros(N, dtype=np.int64); vmem = np.zeros(N, dtype=np.int64)
    prev = np.zeros(N, dtype=np.int64)
    raster = np.zeros((T, N), dtype=np.uint8)
    w = w_in.astype(np.int64)

    def decay(v, dash):
        dv = v >> dash
        dv = np.where(dv == 0, np.sign(v), dv)
        return v - dv
    for t in range(T):
        isyn = decay(isyn, dash_syn) + spikes_in[t].astype(np.int64) @ w
        if w_rec is not None:
            isyn = isyn + prev @ w_rec.astype(np.int64)
        isyn = np.clip(isyn, -32768, 32767)
        vmem = np.clip(decay(vmem, dash_mem) + isyn, -32768, 32767)
        ns = np.clip(vmem // thr, 0, max_spikes)
        ns = np.where(vmem >= thr, ns, 0)
        vmem = vmem - ns * thr
        raster[t] = ns
        prev = ns
    return raster


@pytest.mark.parametrize("name", H.XYLO_CASES)
def test_lif_restatement_equals_independent_numpy_version(name):
    g = H.load(name)
    net = H.xylo_network(g)
    cfg = H.xylo_oracle_cfg(g, net)
    raster, counts = O.xylo_lif(cfg, g["spikes_in"])
    ref = numpy_lif(g["spikes_in"], net.w_in, net.threshold.astype(np.int64), net.dash_syn.astype(np.int64),
                    net.dash_mem.astype(np.int64))
    assert np.array_equal(raster, ref)
    assert np.array_equal(counts, ref.sum(axis=0))
    assert counts.sum() > 0, "the network should fire on a 10 dB clip"


def test_lif_with_recurrent_weights_and_small_thresholds():
    rng = np.random.default_rng(3)
    g = H.load("xylo_3band_o2")
    net = H.xylo_network(g)
    N = net.w_in.shape[1]
    net.w_rec = rng.integers(-3, 2, size=(N, N)).astype(np.int8)
    net.threshold = rng.integers(5, 60, size=N).astype(np.int16)      # multi-spike steps, saturation of the count
    net.dash_syn = rng.integers(0, 5, size=N).astype(np.int8)
    net.dash_mem = rng.integers(1, 6, size=N).astype(np.int8)
    cfg = H.xylo_oracle_cfg(g, net)
    s = g["spikes_in"][:600]
    raster, counts = O.xylo_lif(cfg, s)
    ref = numpy_lif(s, net.w_in, net.threshold.astype(np.int64), net.dash_syn.astype(np.int64),
                    net.dash_mem.astype(np.int64), w_rec=net.w_rec)
    assert np.array_equal(raster, ref)
    assert raster.max() > 1


def test_batch_driver_matches_single_clip():
    g = H.load("xylo_c3_unipolar")
    cfg = H.xylo_oracle_cfg(g)
    x = H.xylo_synth_clips(g, 3, 1500, seed=4, int16=True)
    out = O.xylo_run_batch(cfg, x, float(g["fs"]), win=15, nthreads=2, want_spikes=True)
    for i in range(3):
        spk, sgn = O.xylo_encode(cfg, x[i].astype(np.float64))
        _, counts = O.xylo_lif(cfg, spk)
        assert np.array_equal(out["spikes_signed"][i], sgn)
        assert np.array_equal(out["counts"][i], counts)
        _, d0, d1 = O.xylo_rate_doa(counts, len(g["doa_list"]), 1, 1500, float(g["fs"]), 15)
        assert out["doa"][i] == d0 and out["doa_peak"][i] == d1
