"""Host-side logic of the drop-in package against the golden fixtures (CPU only)."""
import numpy as np
import pytest
from scipy.signal import butter, sosfilt, lfilter

import helpers as H
from haghighatshoarmuir2024_b200 import array_geometry as AG
from haghighatshoarmuir2024_b200 import distributed as D
from haghighatshoarmuir2024_b200.engine import neuron_alpha_params
from haghighatshoarmuir2024_b200.utils import Envelope, find_peak_location


def test_center_circular_geometry_matches_reference():
    g = H.load("snn_c1_bipolar")
    geo = AG.CenterCircularArray(radius=4.5e-2, num_mic=7)
    np.testing.assert_array_equal(geo.r_vec, g["r_vec"])
    np.testing.assert_array_equal(geo.theta_vec, g["theta_vec"])
    d = geo.delays(0.3, normalized=True)
    assert d.min() == 0.0 and len(geo) == 7
    np.testing.assert_allclose(geo.delays_batch([0.3, 1.0])[0], geo.delays(0.3, normalized=False))


def test_linear_geometry_matches_reference():
    g = H.load("snn_linear16")
    geo = AG.LinearArray(spacing=2 * 4.5e-2 / 16, num_mic=16, radius=4.5e-2)
    np.testing.assert_allclose(geo.r_vec, g["r_vec"], rtol=0, atol=0)
    np.testing.assert_array_equal(geo.theta_vec, g["theta_vec"])


def test_geometry_rejects_negative_radius():
    with pytest.raises(ValueError):
        AG.ArrayGeometry(np.array([-1.0]), np.array([0.0]))


@pytest.mark.parametrize("name", H.SNN_CASES)
def test_neuron_alpha_closed_form_equals_reference_taps(name):
    g = H.load(name)
    T = g["x"].shape[0]
    taps, a, c, L = neuron_alpha_params(np.arange(T) / float(g["fs"]), [float(g["tau"])] * 2)
    assert L == len(g["nir"])
    np.testing.assert_allclose(taps, g["nir"], rtol=1e-13)
    n = np.arange(L)
    np.testing.assert_allclose(c * n * a ** n, g["nir"], rtol=1e-10, atol=1e-18)


def test_neuron_recurrence_equals_truncated_fir():
    """The 4-state recurrence of csrc/micloc_device.cuh:neuron_step, in float64."""
    g = H.load("snn_c1_bipolar")
    nir = g["nir"]; L = len(nir)
    _, a, c, _ = neuron_alpha_params(np.arange(4800) / 48000.0, [float(g["tau"])] * 2)
    s = g["spikes"][:, 3].astype(np.float64)
    ref = lfilter(nir, [1], s)
    p1 = p2 = q1 = q2 = 0.0
    out = np.empty_like(s)
    for t in range(len(s)):
        sd = s[t - L] if t >= L else 0.0
        p2 = a * (p2 + p1); p1 = a * p1 + s[t]
        q2 = a * (q2 + q1); q1 = a * q1 + sd
        out[t] = c * p2 - c * a ** L * (q2 + L * q1)
    np.testing.assert_allclose(out, ref, rtol=0, atol=1e-12)


def test_sos_cascade_equals_ba_filter():
    g = H.load("snn_band2_sine")
    sos = butter(2, g["band"], btype="bandpass", output="sos", fs=float(g["fs"]))
    x = np.random.default_rng(0).standard_normal(3000)
    np.testing.assert_allclose(sosfilt(sos, x), lfilter(g["ba_b"], g["ba_a"], x), rtol=0, atol=1e-9)


def test_find_peak_location_matches_reference():
    g = H.load("utils")
    for s, row in zip(g["sigs"], g["idx"]):
        for w, want in zip(g["wins"], row):
            assert find_peak_location(s, int(w)) == int(want)
    with pytest.raises(ValueError):
        find_peak_location(np.zeros((2, 2)), 3)
    with pytest.raises(ValueError):
        find_peak_location(np.zeros(10), 4)
    with pytest.raises(ValueError):
        find_peak_location(np.zeros(10), 7)


def test_envelope_rise_fall():
    env = Envelope(rise_time=1e-3, fall_time=1e-2, fs=48000)
    x = np.concatenate([np.zeros((10, 1)), np.ones((200, 1)), np.zeros((200, 1))])
    y = env.evolve_host(x)
    assert y.shape == x.shape and y[100, 0] > 0.8 and 0 < y[-1, 0] < y[209, 0]
    g = H.load("envelope")                  # outputs of the reference's own Envelope.evolve
    for i in range(int(g["n_cases"])):
        e = Envelope(float(g[f"rise_{i}"]), float(g[f"fall_{i}"]), float(g["fs"]))
        np.testing.assert_allclose(e.evolve_host(g[f"x_{i}"]), g[f"env_{i}"], rtol=1e-12, atol=0)
    with pytest.raises(ValueError):
        Envelope(rise_time=1.0, fall_time=0.1, fs=48000)


def test_shard_range_covers_everything_once():
    for total in (0, 1, 7, 100_000):
        for world in (1, 2, 3, 8):
            spans = [D.shard_range(total, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == total
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        D.shard_range(10, 2, 2)


def test_doa_histogram_groups():
    import torch
    doa = torch.tensor([0, 3, 3, 1, 0], dtype=torch.int32)
    grp = torch.tensor([0, 0, 1, 1, 1])
    h = D.doa_histogram(doa, 4, grp, 2)
    assert h.tolist() == [[1, 0, 0, 1], [1, 1, 0, 1]]
    assert D.doa_histogram(doa, 4).tolist() == [[2, 1, 0, 2]]


@pytest.mark.parametrize("tap_first,n_taps,T", [(1, 240, 1500), (3, 24, 700), (0, 8, 300), (1, 17, 401)])
def test_polyphase_toeplitz_block_algebra_of_the_tensor_core_stht(tap_first, n_taps, T):
    """The index algebra of k_stht_tc (csrc/micloc_staged.cuh), restated in numpy: a stride-2 FIR as two dense FIRs over
    the input parities, each a product of 16 x 16 Toeplitz blocks that depend on (row block - k step) only:
        Q[2u + pi] = sum_j g[j] xs[u - j - c],  xs[v] = x[2v + rho],  pi = (rho + tap_first) & 1,  c = (tap_first + rho - pi) / 2,
        tile: D[a][n] = sum_e A[a][e] B[e][n],  A[a][e] = g[a - e + L],  L = 16 (ND - 1),  ND = (n_taps + 14) // 16 + 1,
        block (mb, ks): tap index j = 16 d + r - c',  d = mb - ks + ND - 1  (0 <= d < ND, else the block is zero).
    Checked against np.convolve with the zero-stuffed kernel (zero initial state: lfilter's)."""
    rng = np.random.default_rng(n_taps)
    g = rng.standard_normal(n_taps)
    x = rng.standard_normal(T)
    h = np.zeros(tap_first + 2 * (n_taps - 1) + 1)
    h[tap_first::2] = g
    ref = np.convolve(x, h)[:T]

    MB = 8                                            # row blocks per parity and tile (kStMB)
    U = 16 * MB
    ND = (n_taps + 14) // 16 + 1
    L = 16 * (ND - 1)
    W = 16 * (MB + ND - 1)
    frag = np.zeros((ND, 16, 16))                     # the ND distinct blocks (kernel: per-lane register fragments)
    for d in range(ND):
        for r in range(16):
            for c in range(16):
                j = 16 * d + r - c
                if 0 <= j < n_taps:
                    frag[d, r, c] = g[j]
    out = np.zeros(T)
    for u0 in range(0, (T + 1) // 2 + U, U):
        for rho in (0, 1):
            pi = (rho + tap_first) & 1
            c = (tap_first + rho - pi) // 2
            vb = u0 - c - L
            idx = 2 * (vb + np.arange(W)) + rho        # staged window of the input parity rho
            win = np.where((idx >= 0) & (idx < T), x[np.clip(idx, 0, T - 1)], 0.0)
            for mb in range(MB):
                acc = np.zeros(16)
                for ks in range(mb, mb + ND):          # the ND k steps whose block is not zero
                    acc += frag[mb - ks + ND - 1] @ win[16 * ks: 16 * ks + 16]
                t = 2 * (u0 + 16 * mb + np.arange(16)) + pi
                ok = t < T
                out[t[ok]] = acc[ok]
    assert np.allclose(out, ref, rtol=1e-12, atol=1e-12)
