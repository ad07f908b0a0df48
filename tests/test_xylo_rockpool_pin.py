"""Consumes tests/golden/xylosim_pin_*.npz -- fixtures only a host with rockpool[xylo] can write
(tests/golden/make_golden_xylosim.py; rockpool is absent from the build container).  Without them every test here
SKIPS and the Xylo integer network (quantisation + hidden-layer dynamics) stays PARITY UNPINNED.

With them:
  * quantize_network == the specification rockpool's mapper + global_quantize produced (weights, dash, thresholds),
  * oracle mo_xylo_lif on the reference's own input spikes == XyloSim's rec["Spikes"], bit for bit,
  * (-m gpu) the CUDA chain on the same int16 clip == both.
"""
import glob
import os

import numpy as np
import pytest

import helpers as H

PINS = sorted(glob.glob(os.path.join(H.GOLDEN, "xylosim_pin_*.npz")))
needs_pin = pytest.mark.skipif(not PINS, reason="no tests/golden/xylosim_pin_*.npz: rockpool is not installed here, the "
                               "Xylo integer network is PARITY UNPINNED (run tests/golden/make_golden_xylosim.py "
                               "where rockpool[xylo] is available)")


def _net(g):
    from haghighatshoarmuir2024_b200.xylo_snn_localization import quantize_network
    return quantize_network(list(g["bf_mats"]), g["taus"], float(g["fs"]), bool(g["bipolar"]))


def test_pin_script_is_importable_and_documents_itself():
    """The pin script itself must stay runnable: it compiles, names its cases, and refuses politely without rockpool."""
    path = os.path.join(H.GOLDEN, "make_golden_xylosim.py")
    src = open(path).read()
    compile(src, path, "exec")
    assert "PARITY UNPINNED" in src and "xylosim_pin_" in src


@needs_pin
@pytest.mark.parametrize("path", PINS)
def test_quantisation_matches_rockpool_specification(path):
    g = dict(np.load(path, allow_pickle=False))
    net = _net(g)
    N = net.w_in.shape[1]
    w_in = np.asarray(g["spec_weights_in"]).reshape(net.w_in.shape[0], -1)[:, :N]
    assert np.array_equal(net.w_in.astype(np.int64), w_in.astype(np.int64))
    assert np.array_equal(net.dash_mem.astype(np.int64), np.asarray(g["spec_dash_mem"]).astype(np.int64)[:N])
    assert np.array_equal(net.dash_syn.astype(np.int64), np.asarray(g["spec_dash_syn"]).astype(np.int64)[:N])
    assert np.array_equal(net.threshold.astype(np.int64), np.asarray(g["spec_threshold"]).astype(np.int64)[:N])
    w_rec = np.asarray(g["spec_weights_rec"]).reshape(-1)
    if net.w_rec is None:
        assert not np.any(w_rec[: N * N].reshape(-1))
    for k, v in (("spec_weight_shift_in", net.weight_shift_in), ("spec_weight_shift_rec", net.weight_shift_rec)):
        if k in g:
            assert int(g[k]) == v


@needs_pin
@pytest.mark.parametrize("path", PINS)
def test_oracle_network_matches_xylosim_raster(path):
    from oracle import oracle as O
    g = dict(np.load(path, allow_pickle=False))
    net = _net(g)
    cfg = O.XyloConfig(h=g["kernel"], b=np.zeros((len(g["bands"]), 3)), a=np.ones((len(g["bands"]), 3)),
                       robust_width=float(g["robust_width"]), bipolar=bool(g["bipolar"]), num_mic=g["x"].shape[1],
                       num_doa=len(g["doa_list"]), w_in=net.w_in, threshold=net.threshold, dash_syn=net.dash_syn,
                       dash_mem=net.dash_mem, w_rec=net.w_rec, bias=net.bias, weight_shift_in=net.weight_shift_in,
                       weight_shift_rec=net.weight_shift_rec, max_spikes=net.max_spikes)
    raster, counts = O.xylo_lif(cfg, g["spikes_in"])
    N = raster.shape[1]
    assert np.array_equal(raster, g["raster"][:, :N])
    assert np.array_equal(counts, g["raster"][:, :N].astype(np.int64).sum(0))


@needs_pin
@pytest.mark.gpu
@pytest.mark.parametrize("path", PINS)
def test_cuda_chain_matches_xylosim(path):
    import torch
    from scipy.signal import butter
    from haghighatshoarmuir2024_b200.xylo_snn_localization import XyloEngine
    g = dict(np.load(path, allow_pickle=False))
    net = _net(g)
    fs = float(g["fs"])
    sos = [butter(1, band, btype="bandpass", output="sos", fs=fs) for band in g["bands"]]
    ba = [butter(1, band, btype="bandpass", output="ba", fs=fs) for band in g["bands"]]
    eng = XyloEngine(num_mic=g["x"].shape[1], stht_kernel=g["kernel"], sos_list=sos, ba_list=ba,
                     robust_width=int(np.ceil(float(g["robust_width"]))), bipolar=bool(g["bipolar"]), net=net,
                     num_doa=len(g["doa_list"]), device=0)
    out = eng.run(torch.from_numpy(g["x"]).cuda()[None], exact=True, want_spikes_in=True, want_raster=True)
    N = out["raster"].shape[2]
    assert np.array_equal(out["spikes_in"][0].cpu().numpy(), g["spikes_in"])
    assert np.array_equal(out["raster"][0].cpu().numpy(), g["raster"][:, :N])
