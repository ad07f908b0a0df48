"""ONE-COMMAND PIN of the Xylo integer network against rockpool's XyloSim -- for a maintainer who has rockpool.

rockpool / xylosim / samna are an un-pinned pip dependency of the reference (setup.py:20) and are NOT installed in
the build container (no network), so the integer hidden layer of this repo (`quantize_network`,
oracle `mo_xylo_lif`, CUDA `k_xylo_lif`) is checked only against its own CPU restatement: PARITY UNPINNED.
This script closes that gap wherever rockpool is importable:

    pip install rockpool[xylo]          # the reference's dependency
    python tests/golden/make_golden_xylosim.py /path/to/HaghighatshoarMuir2024
    python -m pytest tests/test_xylo_rockpool_pin.py -q          # CPU: quantisation + oracle;  -m gpu: the CUDA kernel

It builds the UNMODIFIED reference `micloc.xylo_snn_localization.Demo` (only matplotlib, micloc.record and
micloc.visualizer -- recorder / GUI imports with no arithmetic -- are replaced by empty stand-ins), lets it run
mapper -> global_quantize -> config_from_specification -> XyloSim.from_config (xylo_snn_localization.py:268-290), and
dumps, per case, into tests/golden/xylosim_pin_<case>.npz:

    spec_*          every array of the quantised specification handed to config_from_specification
                    (weights_in, weights_rec, dash_mem, dash_syn, threshold, weight_shift_in, weight_shift_rec, bias ...)
    x               the int16 test clip [T, 7]
    spikes_in       Demo.spike_encoding(x)                                   [T, N_in]
    raster          Demo.xylo_process(spikes_in) = rec["Spikes"]              [T, N]
    vmem, isyn      rec["Vmem"], rec["Isyn"] when XyloSim records them        [T, N]
    rate, doa_*     Demo.extract_rate / estimate_doa_from_rate for the three methods

tests/test_xylo_rockpool_pin.py consumes the files when present and skips (saying "parity unpinned") otherwise.
"""
import contextlib
import io
import os
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
FS = 48_000

CASES = {
    # name: (bands, bipolar, G, T, snr_db, seed)            BASELINE configs[2]: bipolar and unipolar RZCC
    "c3_bipolar": ([[1600.0, 2400.0]], True, 449, 12_000, 10.0, 11),
    "c3_unipolar": ([[1600.0, 2400.0]], False, 449, 12_000, 10.0, 12),
    "3band": ([[1600.0, 2000.0], [2000.0, 2300.0], [2300.0, 2600.0]], True, 225, 12_000, 5.0, 13),
}


def main():
    ref = sys.argv[1] if len(sys.argv) > 1 else "/root/reference"
    try:
        import rockpool  # noqa: F401
        from rockpool.devices.xylo.syns61201 import XyloSim  # noqa: F401
    except Exception as e:
        raise SystemExit(f"rockpool[xylo] is not importable here ({type(e).__name__}: {e}); nothing written -- "
                         "the Xylo integer network stays PARITY UNPINNED")
    for m in ("matplotlib", "matplotlib.pyplot", "micloc.record", "micloc.visualizer"):
        sys.modules.setdefault(m, types.ModuleType(m))
    sys.modules["micloc.record"].AudioRecorder = None
    sys.modules["micloc.visualizer"].Visualizer = None
    sys.path.insert(0, ref)
    import rockpool.devices.xylo.syns61201 as xa2
    import micloc.xylo_snn_localization as X
    from micloc.array_geometry import CenterCircularArray

    # the specification is a local of Demo._initialize_snn_module: record what it hands to config_from_specification
    captured = {}
    real_cfs = X.config_from_specification

    def spy(**spec):
        captured.clear()
        captured.update(spec)
        return real_cfs(**spec)

    X.config_from_specification = spy
    geometry = CenterCircularArray(radius=4.5e-2, num_mic=7)
    for name, (bands, bipolar, G, T, snr_db, seed) in CASES.items():
        np.random.seed(seed)
        doa_list = np.linspace(-np.pi, np.pi, G)
        with contextlib.redirect_stdout(io.StringIO()):
            demo = X.Demo(geometry=geometry, freq_bands=np.asarray(bands), doa_list=doa_list, recording_duration=0.25,
                          kernel_duration=10e-3, bipolar_spikes=bipolar, xylosim_version=True, fs=FS)
        t = np.arange(T) / FS
        f_lo, f_hi = bands[0]
        src = np.sin(2 * np.pi * np.cumsum(f_lo + (f_hi - f_lo) * t / t[-1]) / FS)
        doa_true = float(np.random.uniform(-np.pi, np.pi))
        sig = X.signal_from_template(geometry, (t, src, doa_true))
        snr = 10 ** ((snr_db - 10 * np.log10((FS / 2) / (f_hi - f_lo))) / 10)
        sig = sig + np.sqrt(np.mean(sig ** 2) / snr) * np.random.randn(*sig.shape)
        x = np.round(sig / np.abs(sig).max() * 12000).astype(np.int16)       # integers: float64 and float32 paths see the same samples
        spikes_in = demo.spike_encoding(x.astype(np.float64))
        demo.xylo.reset_state()
        _, _, rec = demo.xylo(spikes_in, record=True)
        raster = np.asarray(rec["Spikes"])
        out = {"bands": np.asarray(bands), "bipolar": bipolar, "fs": FS, "doa_list": doa_list, "doa_true": doa_true,
               "x": x, "spikes_in": spikes_in.astype(np.int8), "raster": raster.astype(np.uint8),
               "bf_mats": np.asarray(demo.bf_mats), "taus": np.asarray(demo.tau_vecs),
               "kernel": demo.beamfs[0].kernel, "robust_width": demo.beamfs[0].spk_encoder.robust_width,
               "rockpool_version": str(getattr(rockpool, "__version__", "?"))}
        for k in ("Vmem", "Isyn"):
            if k in rec:
                out[k.lower()] = np.asarray(rec[k]).astype(np.int32)
        for k, v in captured.items():
            if v is None:
                continue
            try:
                out["spec_" + k] = np.asarray(v)
            except Exception:
                pass
        rate = demo.extract_rate(raster)
        out["rate"] = rate
        for method in ("peak", "periodic_ml", "trimmed_periodic_ml"):
            try:
                out["doa_" + method] = float(demo.estimate_doa_from_rate(rate, method))
            except Exception:                                     # the reference's trimmed window can raise IndexError
                out["doa_" + method] = np.nan
        path = os.path.join(HERE, f"xylosim_pin_{name}.npz")
        np.savez_compressed(path, **out)
        print(f"wrote {path}: N_in {spikes_in.shape[1]}, N {raster.shape[1]}, {int(raster.sum())} hidden spikes, "
              f"spec keys {sorted(k for k in captured)}")
    X.config_from_specification = real_cfs


if __name__ == "__main__":
    main()
