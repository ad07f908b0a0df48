"""GPU parity tests of the Xylo integer chain (BASELINE config 3), through the C-ABI.

Bar: bit-exact.  With the float64 front end (exact=True) the input spikes equal the reference's
own Demo.spike_encoding output (golden fixtures), and every integer after them (hidden raster,
counts, DoA) equals the CPU oracle's restatement of XyloSim.  The float32 front end is held to
the float path's tolerance (>= 99.9 % spike agreement)."""
import numpy as np
import pytest
import torch

import helpers as H
from haghighatshoarmuir2024_b200 import _native as N
from oracle import oracle as O

pytestmark = pytest.mark.gpu


def dev(x):
    return torch.from_numpy(np.ascontiguousarray(x)).cuda()


@pytest.mark.parametrize("name", H.XYLO_CASES)
def test_exact_chain_matches_reference_spikes_and_oracle_network(name):
    g = H.load(name)
    net = H.xylo_network(g)
    eng = H.xylo_engine(g, net)
    cfg = H.xylo_oracle_cfg(g, net)
    win = int(g["post_win"])
    out = eng.run(dev(g["x"]), exact=True, want_spikes_in=True, want_raster=True, peak_win=win)
    spikes_in = out["spikes_in"][0].cpu().numpy()
    assert np.array_equal(spikes_in, g["spikes_in"]), "input spikes differ from the reference's spike_encoding"
    raster, counts = O.xylo_lif(cfg, g["spikes_in"])
    assert np.array_equal(out["raster"][0].cpu().numpy(), raster)
    assert np.array_equal(out["counts"][0].cpu().numpy(), counts)
    _, d0, d1 = O.xylo_rate_doa(counts, len(g["doa_list"]), len(g["bands"]), g["x"].shape[0], float(g["fs"]), win)
    assert int(out["doa"][0]) == d0 and int(out["doa_peak"][0]) == d1
    assert int(out["flags"].sum()) == 0


@pytest.mark.parametrize("name", H.XYLO_CASES)
def test_fast_front_end_within_float_tolerance(name):
    g = H.load(name)
    eng = H.xylo_engine(g)
    out = eng.run(dev(g["x"]), exact=False, want_spikes_in=True)
    s = out["spikes_in"][0].cpu().numpy()
    agree = 1.0 - np.count_nonzero(s != g["spikes_in"]) / max(int(g["spikes_in"].sum()), 1)
    assert agree >= 0.999, agree


def test_fast_front_end_fused_and_staged_agree(monkeypatch):
    """exact=False uses the fused SNN kernel as front end when it can (one band, <= 8 microphones);
    the staged kernels are the fallback.  Both are float32: same tolerance, near-identical spikes."""
    g = H.load("xylo_c3_bipolar")
    eng = H.xylo_engine(g)
    x = dev(H.xylo_synth_clips(g, 10, 3000, seed=8))
    a = eng.run(x, exact=False, want_spikes_in=True)
    monkeypatch.setenv("MICLOC_XYLO_STAGED_FRONT", "1")
    b = eng.run(x, exact=False, want_spikes_in=True)
    sa, sb = a["spikes_in"].cpu().numpy(), b["spikes_in"].cpu().numpy()
    assert 1.0 - np.count_nonzero(sa != sb) / max(int(sb.sum()), 1) >= 0.999
    assert float((a["doa"] == b["doa"]).float().mean()) >= 0.9


@pytest.mark.parametrize("name", H.XYLO_CASES)
def test_process_only_matches_oracle(name):
    g = H.load(name)
    net = H.xylo_network(g)
    eng = H.xylo_engine(g, net)
    raster, counts = O.xylo_lif(H.xylo_oracle_cfg(g, net), g["spikes_in"])
    out = eng.process(dev(g["spikes_in"]), want_raster=True)
    assert np.array_equal(out["raster"][0].cpu().numpy(), raster)
    assert np.array_equal(out["counts"][0].cpu().numpy(), counts)


def test_multi_spike_steps_saturation_and_bias():
    """Low thresholds (several spikes per step, the 31-spike cap), mixed dashes, a bias, int16 saturation."""
    rng = np.random.default_rng(11)
    g = H.load("xylo_c3_bipolar")
    net = H.xylo_network(g)
    n = net.w_in.shape[1]
    net.threshold = rng.integers(1, 40, size=n).astype(np.int16)
    net.dash_syn = rng.integers(0, 6, size=n).astype(np.int8)
    net.dash_mem = rng.integers(0, 8, size=n).astype(np.int8)
    net.bias = rng.integers(-3, 4, size=n).astype(np.int16)
    net.weight_shift_in = 3
    net.max_spikes = 7
    eng = H.xylo_engine(g, net)
    s = (rng.random((2, 900, net.w_in.shape[0])) < 0.2).astype(np.int8)
    out = eng.process(dev(s), want_raster=True)
    cfg = H.xylo_oracle_cfg(g, net)
    for b in range(2):
        raster, counts = O.xylo_lif(cfg, s[b])
        assert raster.max() == 7
        assert np.array_equal(out["raster"][b].cpu().numpy(), raster)
        assert np.array_equal(out["counts"][b].cpu().numpy(), counts)


@pytest.mark.parametrize("name,int16", [("xylo_c3_bipolar", True), ("xylo_c3_unipolar", False), ("xylo_3band_o2", False)])
def test_batch_exact_equals_oracle(name, int16):
    g = H.load(name)
    net = H.xylo_network(g)
    eng = H.xylo_engine(g, net)
    x = H.xylo_synth_clips(g, 12, 2000, seed=31, int16=int16)
    win = 2 * ((len(g["doa_list"]) // 32) // 2) + 1
    ref = O.xylo_run_batch(H.xylo_oracle_cfg(g, net), x, float(g["fs"]), win=win, nthreads=4, want_spikes=True)
    out = eng.run(dev(x), exact=True, want_spikes_in=True, peak_win=win)
    s = out["spikes_in"].cpu().numpy().astype(np.int8)
    CT = ref["spikes_signed"].shape[2]
    sgn = s[..., :CT] - s[..., CT:] if bool(g["bipolar"]) else s
    assert np.array_equal(sgn, ref["spikes_signed"])
    assert np.array_equal(out["counts"].cpu().numpy(), ref["counts"])
    assert np.array_equal(out["doa"].cpu().numpy(), ref["doa"])
    assert np.array_equal(out["doa_peak"].cpu().numpy(), ref["doa_peak"])
    # float32 front end on the same clips: float-path tolerance on spikes, DoA mostly equal
    fast = eng.run(dev(x), exact=False, want_spikes_in=True)
    sf = fast["spikes_in"].cpu().numpy()
    agree = 1.0 - np.count_nonzero(sf != s) / max(int(s.sum()), 1)
    assert agree >= 0.999, agree


@pytest.mark.parametrize("name,int16", [("xylo_c3_bipolar", True), ("xylo_3band_o2", False)])
def test_fast_exact_front_end_equals_staged_float64(name, int16, monkeypatch):
    """The fast exact front end (register-blocked float64 STHT over the non-zero taps + one warp per clip for band
    filters, cumsum and streaming find_peaks) against the staged float64 kernels (dense FIR in lfilter's order,
    unbounded candidate clusters), on whole seconds."""
    g = H.load(name)
    eng = H.xylo_engine(g)
    x = dev(H.xylo_synth_clips(g, 5, 48_000, seed=77, int16=int16))
    a = eng.run(x, exact=True, want_spikes_in=True)
    monkeypatch.setenv("MICLOC_XYLO_STAGED_F64", "1")
    b = eng.run(x, exact=True, want_spikes_in=True)
    monkeypatch.setenv("MICLOC_XYLO_DENSE_F64", "1")          # and the dense 480-tap FIR in lfilter's order
    c = eng.run(x[:2], exact=True, want_spikes_in=True)
    assert torch.equal(a["spikes_in"], b["spikes_in"]) and torch.equal(a["counts"], b["counts"])
    assert torch.equal(a["spikes_in"][:2], c["spikes_in"]) and torch.equal(a["counts"][:2], c["counts"])
    assert int(a["flags"].sum()) == 0 and int(b["flags"].sum()) == 0


def test_long_digital_silence_falls_back_to_the_unbounded_encoder():
    """Hundreds of milliseconds of exact zeros let the float64 band filter underflow to exact 0.0: flat tops longer
    than the streaming encoder follows.  The library redoes such clips with the unbounded kernels by itself: the
    result is still bit-exact and no flag is left set."""
    g = H.load("xylo_c3_bipolar")
    net = H.xylo_network(g)
    eng = H.xylo_engine(g, net)
    x = H.xylo_synth_clips(g, 3, 48_000, seed=12, int16=True)
    x[1, 1500:44_000] = 0
    x[2, :] = 0
    ref = O.xylo_run_batch(H.xylo_oracle_cfg(g, net), x, float(g["fs"]), win=15, nthreads=3, want_spikes=True)
    out = eng.run(dev(x), exact=True, want_spikes_in=True, peak_win=15)
    s = out["spikes_in"].cpu().numpy().astype(np.int8)
    assert np.array_equal(s[..., :14] - s[..., 14:], ref["spikes_signed"])
    assert np.array_equal(out["counts"].cpu().numpy(), ref["counts"])
    assert np.array_equal(out["doa"].cpu().numpy(), ref["doa"])
    assert int(out["flags"].sum()) == 0


def test_full_size_clip_config3():
    """T = 48 000 (1 s) bipolar, G = 449: exact chain == oracle on whole clips."""
    g = H.load("xylo_c3_bipolar")
    net = H.xylo_network(g)
    eng = H.xylo_engine(g, net)
    x = H.xylo_synth_clips(g, 2, 48_000, seed=5, int16=True)
    ref = O.xylo_run_batch(H.xylo_oracle_cfg(g, net), x, float(g["fs"]), win=15, nthreads=2, want_spikes=True)
    out = eng.run(dev(x), exact=True, want_spikes_in=True, peak_win=15)
    s = out["spikes_in"].cpu().numpy().astype(np.int8)
    assert np.array_equal(s[..., :14] - s[..., 14:], ref["spikes_signed"])
    assert np.array_equal(out["counts"].cpu().numpy(), ref["counts"])
    assert np.array_equal(out["doa"].cpu().numpy(), ref["doa"])
    assert np.array_equal(out["doa_peak"].cpu().numpy(), ref["doa_peak"])
    # the raster of the network-only entry point sums to the counts of the full chain
    o2 = eng.process(out["spikes_in"], want_raster=True)
    assert np.array_equal(o2["raster"].sum(dim=1, dtype=torch.int32).cpu().numpy(), ref["counts"])


def test_ragged_and_tiny_clips():
    g = H.load("xylo_c3_unipolar")
    net = H.xylo_network(g)
    eng = H.xylo_engine(g, net)
    cfg = H.xylo_oracle_cfg(g, net)
    for T in (1, 2, 3, 255, 257, 700):
        x = H.xylo_synth_clips(g, 2, max(T, 4), seed=T, int16=True)[:, :T]
        x = np.ascontiguousarray(x)
        ref = O.xylo_run_batch(cfg, x, float(g["fs"]), win=0, nthreads=1, want_spikes=True)
        out = eng.run(dev(x), exact=True, want_spikes_in=True)
        assert np.array_equal(out["spikes_in"].cpu().numpy(), ref["spikes_signed"]), T
        assert np.array_equal(out["counts"].cpu().numpy(), ref["counts"]), T
        assert np.array_equal(out["doa"].cpu().numpy(), ref["doa"]), T


def test_errors():
    g = H.load("xylo_c3_unipolar")
    net = H.xylo_network(g)
    eng = H.xylo_engine(g, net)
    with pytest.raises(ValueError):
        eng.run(torch.zeros((1, 100, 6), device="cuda"))
    with pytest.raises(ValueError):
        eng.run(torch.zeros((1, 100, 7), device="cuda", dtype=torch.float64))
    with pytest.raises(ValueError):
        eng.run(torch.zeros((1, 100, 7), device="cuda"), peak_win=4)          # even window (utils.py:103)
    with pytest.raises(ValueError):
        eng.process(torch.zeros((1, 100, 5), device="cuda", dtype=torch.int8))
    bad = H.xylo_network(g)
    bad.w_rec = np.ones((bad.w_in.shape[1],) * 2, dtype=np.int8)
    with pytest.raises(N.MiclocError):
        H.xylo_engine(g, bad)


def test_demo_dropin_methods():
    """Demo.spike_encoding / xylo_process / extract_rate with the reference's signatures."""
    from haghighatshoarmuir2024_b200.array_geometry import ArrayGeometry
    from haghighatshoarmuir2024_b200.xylo_snn_localization import Demo
    g = H.load("xylo_c3_bipolar")
    geo = ArrayGeometry(g["r_vec"], g["theta_vec"])
    demo = Demo(geometry=geo, freq_bands=g["bands"], doa_list=g["doa_list"], recording_duration=0.05,
                kernel_duration=float(g["kernel_duration"]), bipolar_spikes=True, fs=float(g["fs"]),
                bf_mats=list(g["bf_mats"]))
    spikes_in = demo.spike_encoding(g["x"])
    assert spikes_in.dtype == np.int64 and np.array_equal(spikes_in, g["spikes_in"])
    raster = demo.xylo_process(spikes_in)
    ref_raster, _ = O.xylo_lif(H.xylo_oracle_cfg(g, demo.net), g["spikes_in"])
    assert np.array_equal(raster, ref_raster)
    rate = demo.extract_rate(raster)
    assert rate.shape == (len(g["doa_list"]),)
    assert demo.estimate_doa_from_rate(rate, "peak") == g["doa_list"][np.argmax(rate)]


@pytest.mark.parametrize("scale,bias", [(1, False), (8, False), (1, True)])
def test_tensor_core_input_kernel_equals_event_loop_kernel_and_oracle(scale, bias, monkeypatch):
    """k_xylo_lif_mma (weighted input by int8 MMA, clamps dropped where the host proves them idle) against the
    event-loop kernel and the oracle: small weights (no clamp can trigger: SAT_ISYN = SAT_V = false), the reference's
    weights, with and without a bias; dense and sparse input, a ragged clip length."""
    rng = np.random.default_rng(5 + scale)
    g = H.load("xylo_c3_bipolar")
    net = H.xylo_network(g)
    n = net.w_in.shape[1]
    net.w_in = np.clip(net.w_in.astype(np.int32) // scale, -127, 127).astype(np.int8)
    net.threshold = np.full(n, max(int(net.threshold[0]) // scale, 2), dtype=np.int16)
    if bias:
        net.bias = rng.integers(-2, 3, size=n).astype(np.int16)
    eng = H.xylo_engine(g, net)
    cfg = H.xylo_oracle_cfg(g, net)
    for T, density in ((1000, 0.04), (253, 0.5)):
        s = (rng.random((3, T, net.w_in.shape[0])) < density).astype(np.int8)
        monkeypatch.delenv("MICLOC_XYLO_LIF_ADDS", raising=False)
        mma = eng.process(dev(s), want_raster=True)
        monkeypatch.setenv("MICLOC_XYLO_LIF_ADDS", "1")
        adds = eng.process(dev(s), want_raster=True)
        assert torch.equal(mma["raster"], adds["raster"]) and torch.equal(mma["counts"], adds["counts"])
        raster, counts = O.xylo_lif(cfg, s[0])
        assert np.array_equal(mma["raster"][0].cpu().numpy(), raster)
        assert np.array_equal(mma["counts"][0].cpu().numpy(), counts)
