#!/usr/bin/env python
"""Time the UNMODIFIED reference (micloc.snn_beamformer.SNNBeamformer.apply_to_signal + the caller's power / argmax,
paper_plots/target_snn_localization.py:447-467) on bench.py's workload, in a multiprocessing.Pool over the host cores
(BASELINE.md section 3), beside the C oracle port on the same clips and cores.

Needs /root/reference (build container only -- the GPU box does not have it); writes
profiles/r02_reference_python_timing.json, which `bench.py --impl reference` attaches to its line as
cpu_baseline.reference_python when the reference cannot be imported at run time.

    python tools/time_reference_python.py [clips_per_band]
"""
import json
import multiprocessing as mp
import os
import sys
import time
import types

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference"


def _import_reference():
    for m in ("matplotlib", "matplotlib.pyplot", "matplotlib.gridspec"):
        sys.modules.setdefault(m, types.ModuleType(m))
    sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]
    if REF not in sys.path:
        sys.path.insert(0, REF)
    from micloc.array_geometry import CenterCircularArray
    from micloc.snn_beamformer import SNNBeamformer
    return CenterCircularArray, SNNBeamformer


_STATE = {}


def _worker_init(bands, taus, kernel_duration, fs):
    import contextlib
    import io
    CenterCircularArray, SNNBeamformer = _import_reference()
    geo = CenterCircularArray(radius=4.5e-2, num_mic=7)
    with contextlib.redirect_stdout(io.StringIO()):
        _STATE["bf"] = [SNNBeamformer(geo, kernel_duration, band, [tau, tau], bipolar_spikes=True, fs=fs)
                        for band, tau in zip(bands, taus)]


def _worker(job):
    import numpy as np
    band_idx, bf_mat, t, x = job
    y = _STATE["bf"][band_idx].apply_to_signal(bf_mat, (t, x))
    return int(np.argmax(np.mean(np.abs(y) ** 2, axis=0)))


def time_reference(n_per_band, cores=None):
    """(clips/s of the reference Python on `cores` processes, DoA lists, clips, seconds)."""
    import numpy as np
    sys.path.insert(0, ROOT)
    import bench as Bn
    d, bands = Bn.load_workload()
    cores = cores or os.cpu_count() or 1
    taus = [float(d[f"tau_{i}"]) for i in range(len(bands))]
    clips = [Bn.host_clips(d, i, bands, n_per_band, seed=1000 + i) for i in range(len(bands))]
    t = np.arange(Bn.T_CLIP) / Bn.FS
    jobs = [(i, d[f"bf_{i}"], t, clips[i][j].astype(np.float64)) for i in range(len(bands)) for j in range(n_per_band)]
    with mp.Pool(cores, initializer=_worker_init, initargs=(bands, taus, float(d["kernel_duration"]), Bn.FS)) as pool:
        pool.map(_worker, jobs[:cores])                           # warm-up: imports, filter design, page-in
        t0 = time.perf_counter()
        doa = pool.map(_worker, jobs, chunksize=1)
        secs = time.perf_counter() - t0
    return len(jobs) / secs, doa, clips, secs, cores


def main():
    import numpy as np
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 8
    if not os.path.isdir(REF):
        raise SystemExit(f"{REF} is not here: the reference Python can only be timed in the build container")
    v, doa, clips, secs, cores = time_reference(n)
    import bench as Bn
    d, bands = Bn.load_workload()
    cfgs = Bn.oracle_cfgs(d, bands)
    Bn.cpu_pass(cfgs, [c[:1] for c in clips], cores)
    n_port, s_port, outs = Bn.cpu_pass(cfgs, clips, cores)
    port_doa = np.concatenate([o["doa"] for o in outs])
    res = {
        "what": "unmodified micloc.snn_beamformer.SNNBeamformer.apply_to_signal + mean |y|^2 + argmax per clip "
                "(paper_plots/target_snn_localization.py:447-467) in a multiprocessing.Pool, bench.py's configs[1] clips",
        "where": "build container (no GPU); the GPU box has no /root/reference",
        "cores": cores, "clips": len(doa), "seconds": secs, "reference_python_clips_per_s": v,
        "oracle_port_clips_per_s_same_host": n_port / s_port,
        "port_over_python": (n_port / s_port) / v,
        "doa_identical_port_vs_python": bool(np.array_equal(port_doa, np.asarray(doa))),
    }
    out = os.path.join(ROOT, "profiles", "r02_reference_python_timing.json")
    json.dump(res, open(out, "w"), indent=1)
    print(json.dumps(res, indent=1))


if __name__ == "__main__":
    main()
