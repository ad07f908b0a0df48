"""Multi-band localiser (filterbank -> one SNN chain per band -> power summed over bands -> estimators) against a
golden frame processed by the reference's own classes with the semantics of micloc/localization_demo_snn.py:125-193
(tests/golden/make_golden_full.py:multiband_case)."""
import ctypes as C

import numpy as np
import pytest
import torch
from scipy.signal import butter, lfilter

import helpers as H

pytestmark = pytest.mark.gpu


def make_demo(g):
    from haghighatshoarmuir2024_b200.array_geometry import ArrayGeometry
    from haghighatshoarmuir2024_b200.localization_demo_snn import Demo
    geo = ArrayGeometry(g["r_vec"], g["theta_vec"])
    return Demo(geometry=geo, freq_bands=g["bands"], doa_list=g["doa_list"], recording_duration=0.25,
                kernel_duration=float(g["kernel_duration"]), bipolar_spikes=bool(g["bipolar"]), fs=float(g["fs"]),
                bf_mats=list(g["bf_mats"]))


def test_frame_matches_reference_golden():
    g = H.load("full_multiband")
    demo = make_demo(g)
    frame = torch.from_numpy(g["frame"]).cuda()                       # int32 [T, 8], last channel is not a microphone
    out = demo.localize(frame, want_spikes=True)
    torch.cuda.synchronize()
    assert int(out["flags"][0]) == 0
    # filterbank rows against scipy on the same integers
    x = g["frame"][:, :-1].astype(np.float64)
    for f in range(3):
        ref = lfilter(g["fb_b"][f], g["fb_a"][f], x, axis=0)
        assert H.rel_err(out["banded"][f, 0].cpu().numpy(), ref) < 1e-4
        assert H.spike_agreement(out["spikes"][f][0].cpu().numpy(), g["spikes"][f]) >= 0.999
        assert H.rel_err(out["power"][f, 0].cpu().numpy(), g["powers"][f]) < 2e-3
    assert H.rel_err(out["power_grid"][0].cpu().numpy(), g["power_grid"]) < 2e-3
    assert int(out["doa"][0]) == int(g["doa"])
    assert bool(out["active"][0])
    assert abs(float(out["doa_deg"][0]) - float(g["doa_list"][int(g["doa"])]) * 180 / np.pi) < 1e-9
    wrap = lambda a: (a + np.pi) % (2 * np.pi) - np.pi
    assert abs(wrap(float(out["periodic_ml"][0]) - float(g["periodic_ml"]))) < 2e-3
    assert abs(wrap(float(out["trimmed_periodic_ml"][0]) - float(g["trimmed_periodic_ml"]))) < 2e-3
    # process_frame = the body of the reference's loop
    assert abs(demo.process_frame(g["frame"]) - float(g["doa_list"][int(g["doa"])]) * 180 / np.pi) < 1e-9


def test_silent_frame_is_reported_inactive_and_batches_work():
    g = H.load("full_multiband")
    demo = make_demo(g)
    frames = np.stack([g["frame"], np.zeros_like(g["frame"]), g["frame"] // 4])
    out = demo.localize(torch.from_numpy(frames).cuda())
    assert out["active"].cpu().tolist() == [True, False, True]
    deg = out["doa_deg"].cpu().numpy()
    assert np.isnan(deg[1]) and deg[0] == deg[2]
    assert int(out["doa"][0]) == int(g["doa"]) == int(out["doa"][2])
    with pytest.raises(ValueError):
        demo.localize(torch.zeros((100, 6), dtype=torch.int32, device="cuda"))      # fewer channels than microphones


def test_power_fuse_matches_numpy_estimators():
    """micloc_power_fuse against the reference's estimator formulas (xylo_snn_localization.py:424-444) on random
    patterns, including the peak positions where the reference's own indexing raises IndexError."""
    from haghighatshoarmuir2024_b200 import _native as N
    lib = N.lib()
    rng = np.random.default_rng(5)
    for G in (64, 449, 33):
        F, B = 3, 40
        doa_list = np.linspace(-np.pi, np.pi, G)
        p = rng.random((F, B, G)).astype(np.float32)
        for b in range(B):                                            # a clear peak at a chosen position
            p[:, b, (b * G) // B] += 2.0
        pd = torch.from_numpy(p).cuda()
        dl = torch.from_numpy(doa_list).cuda()
        ps = torch.empty((B, G), dtype=torch.float32, device="cuda")
        doa = torch.empty(B, dtype=torch.int32, device="cuda")
        ml = torch.empty(B, dtype=torch.float64, device="cuda")
        tr = torch.empty(B, dtype=torch.float64, device="cuda")
        fl = torch.empty(B, dtype=torch.int32, device="cuda")
        ptr = lambda t: C.c_void_p(t.data_ptr())
        N.check(lib.micloc_power_fuse(ptr(pd), F, B, G, ptr(dl), ptr(ps), ptr(doa), ptr(ml), ptr(tr), ptr(fl), 0, None))
        torch.cuda.synchronize()
        grid = p.astype(np.float64).sum(0)
        np.testing.assert_allclose(ps.cpu().numpy(), grid, rtol=1e-6)
        n_err = 0
        for b in range(B):
            idx = int(np.argmax(grid[b]))
            assert int(doa[b]) == idx
            assert abs(float(ml[b]) - np.angle(np.mean(grid[b] * np.exp(1j * doa_list)))) < 1e-9
            num = G // 2
            rng_idx = np.arange(-num // 2, num // 2 + 1) - idx
            try:
                want = np.angle(np.mean(grid[b][rng_idx] * np.exp(1j * doa_list[rng_idx])))
                assert abs(float(tr[b]) - want) < 1e-9 and int(fl[b]) == 0
            except IndexError:
                n_err += 1
                assert np.isnan(float(tr[b])) and int(fl[b]) == 2
        assert n_err > 0


def test_filterbank_orders_and_dtypes():
    from haghighatshoarmuir2024_b200 import _native as N
    lib = N.lib()
    rng = np.random.default_rng(2)
    B, T, ch, M = 3, 3000, 8, 7
    bands = [[1600, 2000], [2300, 2600]]
    for order in (1, 2):
        sos = np.ascontiguousarray([butter(order, b, btype="bandpass", output="sos", fs=48000) for b in bands])
        for dtype in (np.int16, np.int32, np.float32):
            x = (rng.standard_normal((B, T, ch)) * 3000).astype(dtype)
            xd = torch.from_numpy(x).cuda()
            out = torch.empty((2, B, T, M), dtype=torch.float32, device="cuda")
            ss = torch.empty(B, dtype=torch.float64, device="cuda")
            code = {np.int16: N.I16, np.int32: N.I32, np.float32: N.F32}[dtype]
            N.check(lib.micloc_filterbank(C.c_void_p(xd.data_ptr()), code, B, T, ch, M, 2, sos.shape[1],
                                          sos.ctypes.data_as(N._dp), C.c_void_p(out.data_ptr()), C.c_void_p(ss.data_ptr()), 0, None))
            torch.cuda.synchronize()
            xf = x[..., :M].astype(np.float64)
            for f, band in enumerate(bands):
                b_, a_ = butter(order, band, btype="bandpass", fs=48000)
                assert H.rel_err(out[f].cpu().numpy(), lfilter(b_, a_, xf, axis=1)) < 1e-4
            np.testing.assert_allclose(ss.cpu().numpy(), (xf ** 2).sum((1, 2)), rtol=1e-5)
