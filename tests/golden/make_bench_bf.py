"""Beamforming matrices of the headline workload (BASELINE.json configs[1]): the
Monte-Carlo SNR sweep of paper_plots/target_snn_localization.py:309-467 on the three
bands of paper_plots/snn_localization_benchmark.py:556-561.

Produced by the UNMODIFIED reference (SNNBeamformer.design_from_template, imported
from /root/reference with an empty matplotlib stub) so that bench.py's GPU arm and
its --impl reference arm both use the reference's own matrices.  Run here:

    python tests/golden/make_bench_bf.py        # -> tests/golden/bench_c2_bf.npz
"""
import contextlib
import io
import os
import sys
import types

for m in ("matplotlib", "matplotlib.pyplot"):
    sys.modules[m] = types.ModuleType(m)
sys.path.insert(0, "/root/reference")

import numpy as np

from micloc.array_geometry import CenterCircularArray
from micloc.snn_beamformer import SNNBeamformer

HERE = os.path.dirname(os.path.abspath(__file__))
FS = 48_000
BANDS = [[1600, 2000], [2000, 2300], [2300, 2600]]
G = 64 * 7 + 1                      # target_snn_localization.py:366


def main():
    geometry = CenterCircularArray(radius=4.5e-2, num_mic=7)
    t = np.arange(0, 1.0, step=1 / FS)
    doa_list = np.linspace(-np.pi, np.pi, G)
    out = dict(fs=FS, bands=np.asarray(BANDS, float), doa_list=doa_list, r_vec=geometry.r_vec,
               theta_vec=geometry.theta_vec, kernel_duration=10e-3)
    for i, band in enumerate(BANDS):
        tau = 1.0 / (2 * np.pi * band[1])          # target_snn_localization.py:330-332 (freq_design = f_high)
        beamf = SNNBeamformer(geometry, 10e-3, band, np.array([tau, tau]), bipolar_spikes=True, fs=FS)
        f_inst = band[0] + (band[1] - band[0]) * (t % t[-1]) / t[-1]      # chirp template, :351-356
        chirp = np.sin(2 * np.pi * np.cumsum(f_inst) / FS)
        with contextlib.redirect_stdout(io.StringIO()), contextlib.redirect_stderr(io.StringIO()):
            bf = beamf.design_from_template((t, chirp), doa_list)
        out[f"bf_{i}"] = bf
        out[f"tau_{i}"] = tau
        out[f"robust_width_{i}"] = beamf.spk_encoder.robust_width
        if i == 0:
            out["kernel"] = beamf.kernel
        print("band", band, "bf", bf.shape, "w", beamf.spk_encoder.robust_width, flush=True)
    np.savez_compressed(os.path.join(HERE, "bench_c2_bf.npz"), **out)


if __name__ == "__main__":
    main()
