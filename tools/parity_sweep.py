#!/usr/bin/env python
"""DoA / spike parity of the fused CUDA path against the CPU oracle on bench.py's workload (BASELINE configs[1]):
`clips` clips per band (default 2002 = 182 per SNR), all 11 SNRs, all three bands, float32 and int16 input.

Runs on the GPU box (the oracle needs ~0.25 core-seconds per clip):

    python tools/parity_sweep.py [clips_per_band] > profiles/r02_parity_sweep.json

North-star bars: identical DoA argmax on >= 99.5 % of clips, spike time+sign agreement >= 99.9 %.
"""
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import bench as Bn
import helpers as H
from haghighatshoarmuir2024_b200.montecarlo import BandSetup, SnrSweep
from oracle import oracle as O


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 2002
    n_spk = 66                                            # clips per band whose full spike rasters are compared (6 per SNR)
    d, bands = Bn.load_workload()
    cores = os.cpu_count() or 1
    setups = [BandSetup(band=bands[i], tau=float(d[f"tau_{i}"]), bf_mat=d[f"bf_{i}"]) for i in range(len(bands))]
    sweep = SnrSweep(setups, d["r_vec"], d["theta_vec"], Bn.FS, float(d["kernel_duration"]), Bn.T_CLIP, device=0)
    cfgs = Bn.oracle_cfgs(d, bands)
    res = {"workload": Bn.workload_config(n, "f32 and i16")["workload"], "clips_per_band": n, "snr_db": list(Bn.SNR_GRID),
           "oracle": "oracle/micloc_oracle.c (float64, pinned bit-identical to the reference on tests/golden)",
           "bands": []}
    t_start = time.time()
    for i, band in enumerate(bands):
        entry = {"band_hz": band}
        for dtype, name in ((torch.float32, "f32"), (torch.int16, "i16")):
            audio, doa_true, snr_idx = sweep.synthesize(i, n, seed=500 + i, snr_db_grid=Bn.SNR_GRID, dtype=dtype)
            out = sweep.run_band(i, audio, want_spikes=False, want_power=True)
            spk = sweep.run_band(i, audio[:n_spk], want_spikes=True, want_power=False)["spikes"].cpu().numpy()
            torch.cuda.synchronize()
            x = audio.cpu().numpy()
            t0 = time.perf_counter()
            ref = O.snn_run_batch(cfgs[i], x, nthreads=cores, want_power=True, want_spikes=False)
            cpu_s = time.perf_counter() - t0
            ref_spk = O.snn_run_batch(cfgs[i], x[:n_spk], nthreads=cores, want_power=False, want_spikes=True)["spikes"]
            doa = out["doa"].cpu().numpy()
            same = doa == ref["doa"]
            # a mismatch by one grid step at near-equal power is a float32-vs-float64 tie; anything else is a bug
            G = sweep.G
            step = np.abs(doa - ref["doa"])
            step = np.minimum(step, G - step)
            per_snr = [float(same[snr_idx == k].mean()) for k in range(len(Bn.SNR_GRID))]
            agree = [H.spike_agreement(spk[j], ref_spk[j]) for j in range(n_spk)]
            perr = float(np.max(np.abs(out["power"].cpu().numpy() - ref["power"]).max(1) / ref["power"].max(1)))
            entry[name] = {"doa_match_rate": float(same.mean()), "doa_match_rate_per_snr": per_snr,
                           "mismatches": int((~same).sum()), "largest_mismatch_grid_steps": int(step.max()),
                           "power_max_rel_err": perr, "spike_agreement_min": float(min(agree)),
                           "spike_agreement_mean": float(np.mean(agree)), "spike_clips": n_spk,
                           "rzcc_overflow_clips": int(out["flags"].sum()), "cpu_clips_per_sec": n / cpu_s, "cpu_cores": cores}
            print(f"band {i} {name}: DoA match {same.mean():.5f} ({int((~same).sum())} of {n}), spikes min "
                  f"{min(agree):.6f}, power rel err {perr:.2e}, {time.time() - t_start:.0f} s", file=sys.stderr)
            del audio, out
        res["bands"].append(entry)
    allr = [e[k]["doa_match_rate"] for e in res["bands"] for k in ("f32", "i16")]
    res["doa_match_rate_min_over_bands"] = min(allr)
    res["clips_total"] = n * len(bands) * 2
    res["pass"] = bool(min(allr) >= 0.995 and min(e[k]["spike_agreement_min"] for e in res["bands"] for k in ("f32", "i16")) >= 0.999)
    print(json.dumps(res, indent=1))


if __name__ == "__main__":
    main()
