"""On-device Monte-Carlo input synthesis (micloc_synth_clips) against the signals the reference itself builds:
goldens captured from SNNBeamformer.apply_to_template and signal_multiple_targets (tests/golden/make_golden_synth.py)."""
import numpy as np
import pytest
import torch

import helpers as H

pytestmark = pytest.mark.gpu
FS = 48_000


def synth(*a, **k):
    from haghighatshoarmuir2024_b200.montecarlo import synthesize_clips
    return synthesize_clips(*a, **k)


@pytest.mark.parametrize("name", ["sine7", "sine7b", "chirp7", "noise7", "chirp16"])
def test_apply_to_template_signal_matches_reference(name):
    g = H.load("synth")
    x_ref = g[f"m0_{name}_x"]
    T = x_ref.shape[0]
    x = synth(g[f"m0_{name}_r"], g[f"m0_{name}_theta"], FS, T, [float(g[f"m0_{name}_doa"])], source=g[f"m0_{name}_src"], mode=0)
    assert x.shape == (1, T, x_ref.shape[1]) and x.dtype == torch.float32
    assert H.rel_err(x[0].cpu().numpy(), x_ref) < 2e-6           # float32 linear interpolation of the same samples


def test_analytic_sine_source_matches_reference():
    g = H.load("synth")
    for name in ("sine7", "sine7b"):
        x_ref = g[f"m0_{name}_x"]
        x = synth(g[f"m0_{name}_r"], g[f"m0_{name}_theta"], FS, x_ref.shape[0], [float(g[f"m0_{name}_doa"])], sine_freq=2000.0, mode=0)
        assert H.rel_err(x[0].cpu().numpy(), x_ref) < 5e-6


@pytest.mark.parametrize("name", ["two", "three"])
def test_multiple_targets_signal_matches_reference(name):
    g = H.load("synth")
    x_ref = g[f"m1_{name}_x"]
    x = synth(g["r7"], g["theta7"], FS, x_ref.shape[0], g[f"m1_{name}_doa"][None, :], source=g[f"m1_{name}_src"],
              gain=g[f"m1_{name}_gain"][None, :], mode=1)
    assert H.rel_err(x[0].cpu().numpy(), x_ref) < 2e-6


def test_noise_statistics_seed_and_int16():
    g = H.load("synth")
    T, B = 4800, 12
    doa = np.linspace(0, 5, B)
    snr = np.full(B, 1.0)                                          # 0 dB: sigma = rms(clean clip)
    clean = synth(g["r7"], g["theta7"], FS, T, doa, sine_freq=2000.0).cpu().numpy().astype(np.float64)
    a = synth(g["r7"], g["theta7"], FS, T, doa, snr_lin=snr, sine_freq=2000.0, seed=7).cpu().numpy().astype(np.float64)
    b = synth(g["r7"], g["theta7"], FS, T, doa, snr_lin=snr, sine_freq=2000.0, seed=7).cpu().numpy().astype(np.float64)
    c = synth(g["r7"], g["theta7"], FS, T, doa, snr_lin=snr, sine_freq=2000.0, seed=8).cpu().numpy().astype(np.float64)
    assert np.array_equal(a, b) and not np.array_equal(a, c)       # counter-based generator: seed decides
    n = a - clean
    rms = np.sqrt(np.mean(clean ** 2, axis=(1, 2)))
    assert np.allclose(n.std(axis=(1, 2)), rms, rtol=0.02)          # sigma = rms / sqrt(snr) (snn_beamformer.py:270-274)
    assert abs(n.mean()) < 0.01 * rms.mean()
    z = n / rms[:, None, None]
    assert abs(np.mean(z ** 3)) < 0.05 and abs(np.mean(z ** 4) - 3.0) < 0.1      # Gaussian moments
    assert abs(np.corrcoef(n[0].ravel(), n[1].ravel())[0, 1]) < 0.02             # clips are independent
    assert abs(np.corrcoef(n[0, :-1].ravel(), n[0, 1:].ravel())[0, 1]) < 0.02    # white
    q = synth(g["r7"], g["theta7"], FS, T, doa, snr_lin=snr, sine_freq=2000.0, seed=7, dtype=torch.int16).cpu().numpy()
    assert q.dtype == np.int16 and np.all(np.abs(q).reshape(B, -1).max(axis=1) == 12000)
    assert np.allclose(q / 12000.0, a / np.abs(a).reshape(B, -1).max(axis=1)[:, None, None], atol=1.0 / 12000)


def test_snr_sweep_synthesize_runs_on_the_kernel():
    import bench as Bn
    from haghighatshoarmuir2024_b200.montecarlo import BandSetup, SnrSweep
    d, bands = Bn.load_workload()
    sweep = SnrSweep([BandSetup(band=bands[0], tau=float(d["tau_0"]), bf_mat=d["bf_0"])], d["r_vec"], d["theta_vec"],
                     Bn.FS, float(d["kernel_duration"]), 4800, device=0)
    x, doa, snr_idx = sweep.synthesize(0, 22, seed=3, snr_db_grid=Bn.SNR_GRID)
    assert x.shape == (22, 4800, 7) and doa.shape == (22,) and snr_idx.max() == 10
    # the quietest clip (20 dB before the bandwidth correction) is dominated by the sine, the noisiest by noise
    p = x.pow(2).mean(dim=(1, 2)).cpu().numpy()
    assert p[0] > 5 * p[10]
    out = sweep.run_band(0, x)
    assert out["doa"].shape == (22,)


def test_bad_arguments_are_rejected():
    g = H.load("synth")
    with pytest.raises(ValueError):
        synth(g["r7"], g["theta7"], FS, 4800, [0.1, 0.2], source=g["m1_two_src"], source_index=[0, 5])
    with pytest.raises(ValueError):
        synth(g["r7"], g["theta7"], FS, 1, [0.1], sine_freq=100.0)
