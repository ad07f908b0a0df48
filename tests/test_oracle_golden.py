"""Pin the CPU oracle (oracle/micloc_oracle.c) against golden vectors produced by the
real reference (tests/golden/make_golden.py).  CPU only."""
import numpy as np
import pytest

from oracle import oracle as O
import helpers as H


@pytest.mark.parametrize("name", H.SNN_CASES)
def test_snn_chain_matches_reference(name):
    g = H.load(name)
    x = g["x"].astype(np.float64)
    out = O.snn_apply(H.oracle_cfg(g), x)
    rows = g["rows"]
    # scipy's FIR path (np.convolve) sums in another order than the DF2T delay line:
    # equal to ~1e-15 of full scale, amplified to ~1e-12 by the band-pass recursion
    assert H.rel_err(out["q"][rows], g["q_rows"]) < 1e-13
    assert H.rel_err(out["z"][rows], g["z_rows"]) < 1e-10
    assert np.array_equal(out["spikes"].astype(np.int8), g["spikes"])          # bit-exact spikes
    assert H.rel_err(out["vmem"][rows], g["vmem_rows"]) < 1e-12
    assert H.rel_err(out["y"][rows], g["y_rows"]) < 1e-12
    assert H.rel_err(out["power"], g["power"]) < 1e-12
    assert out["doa"] == int(g["doa"])


@pytest.mark.parametrize("name", H.FULL_CASES)
def test_snn_chain_matches_reference_at_full_length(name):
    """One-second clips (T = 48 000, configs[0] and one clip per band of configs[1]): the neuron kernel's
    normalisation over T and a whole second of spikes."""
    g = H.load(name)
    assert g["x"].shape[0] == 48_000
    out = O.snn_apply(H.oracle_cfg(g), g["x"].astype(np.float64))
    rows = g["rows"]
    assert H.rel_err(out["q"][rows], g["q_rows"]) < 1e-13
    assert np.array_equal(out["spikes"].astype(np.int8), g["spikes"])          # bit-exact spikes
    assert H.rel_err(out["vmem"][rows], g["vmem_rows"]) < 1e-12
    assert H.rel_err(out["y"][g["yrows"]], g["y_rows"]) < 1e-12
    assert H.rel_err(out["power"], g["power"]) < 1e-12
    assert out["doa"] == int(g["doa"])


@pytest.mark.parametrize("name", H.SNN_CASES + H.FULL_CASES)
def test_neuron_kernel_matches_reference(name):
    g = H.load(name)
    T = g["x"].shape[0]
    t = np.arange(T) / float(g["fs"])
    nir = O.neuron_kernel(t, float(g["tau"]), float(g["tau"]))
    assert len(nir) == len(g["nir"])
    np.testing.assert_allclose(nir, g["nir"], rtol=1e-13, atol=0)


def test_neuron_kernel_rejects_unequal_taus():
    with pytest.raises(AssertionError):
        O.neuron_kernel(np.arange(100) / 48000.0, 1e-4, 2e-4)


def test_rzcc_matches_reference_bit_exact():
    g = H.load("rzcc")
    n = 0
    for key in g:
        if not key.startswith("spk_"):
            continue
        _, sname, w, b = key.split("_")
        got = O.rzcc(g["sig_" + sname], float(w[1:]), bool(int(b[1:])))
        assert np.array_equal(got.astype(np.int8), g[key]), key
        n += 1
    assert n == 32


def test_rzcc_rejects_distance_below_one():
    with pytest.raises(ValueError):
        O.rzcc(np.zeros((10, 1)), 0.5, False)


def test_rzcc_edge_cases():
    assert O.rzcc(np.zeros((0, 2)).reshape(0, 2), 3, True).shape == (0, 2)
    assert not O.rzcc(np.ones((2, 1)), 3, True).any()            # T < 3: no interior sample
    x = np.array([[1.0], [1.0], [-1.0], [-1.0], [1.0], [1.0]])   # cumsum 1 2 1 0 1 2
    s = O.rzcc(x, 1, True)[:, 0]
    assert list(s) == [0, 1, 0, -1, 0, 0]


def test_beamformer_matches_reference():
    g = H.load("beamformer")
    from scipy.signal import butter
    b, a = butter(2, g["band"], btype="bandpass", output="ba", fs=float(g["fs"]))
    import scipy.signal as ss
    K = int(float(g["fs"]) * 10e-3)
    imp = np.zeros(K); imp[0] = 1
    h = np.fft.fftshift(np.imag(ss.hilbert(imp)))
    y = O.beamformer_apply(g["x"].astype(np.float64), h, b, a, g["bf_mat"])
    assert H.rel_err(y[g["rows"]], g["y_rows"]) < 1e-10
    power = np.mean(np.abs(y) ** 2, axis=0)
    assert H.rel_err(power, g["power"]) < 1e-10
    assert int(np.argmax(power)) == int(g["doa"])


def test_batch_driver_matches_single_clip_and_threads():
    g = H.load("snn_c1_bipolar")
    x, _ = H.synth_clips(g, 6, 1200, seed=3)
    cfg = H.oracle_cfg(g)
    one = O.snn_run_batch(cfg, x, nthreads=1, want_spikes=True)
    many = O.snn_run_batch(cfg, x, nthreads=4, want_spikes=True)
    assert np.array_equal(one["doa"], many["doa"]) and np.array_equal(one["spikes"], many["spikes"])
    np.testing.assert_array_equal(one["power"], many["power"])
    ref = O.snn_apply(cfg, x[2].astype(np.float64), want=("power", "spikes"))
    assert ref["doa"] == one["doa"][2]
    np.testing.assert_array_equal(ref["power"], one["power"][2])
