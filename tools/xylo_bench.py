#!/usr/bin/env python
"""Throughput of the Xylo chain (config 3) on the GPU box: fast / exact front end and the integer network alone."""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import helpers as H

B = int(sys.argv[1]) if len(sys.argv) > 1 else 296
g = H.load("xylo_c3_bipolar")
net = H.xylo_network(g)
eng = H.xylo_engine(g, net)
x = torch.from_numpy(H.xylo_synth_clips(g, 8, 48_000, seed=1, int16=False)).cuda()
x = x.repeat((B + 7) // 8, 1, 1)[:B].contiguous()
x += 1e-3 * torch.randn_like(x)


def timed(fn, n=3):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        out = fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n, out


ms_fast, out = timed(lambda: eng.run(x, exact=False, want_spikes_in=True))
print(f"fast  chain: {ms_fast:8.2f} ms for {B} clips -> {B / ms_fast * 1e3:9.0f} clips/s")
spikes = out["spikes_in"]
ms_lif, _ = timed(lambda: eng.process(spikes, want_raster=False))
print(f"LIF only   : {ms_lif:8.2f} ms for {B} clips -> {B / ms_lif * 1e3:9.0f} clips/s  "
      f"(input density {float(spikes.float().mean()):.4f})")
ms_exact, out2 = timed(lambda: eng.run(x, exact=True), n=1)
print(f"exact chain: {ms_exact:8.2f} ms for {B} clips -> {B / ms_exact * 1e3:9.0f} clips/s")
print("doa agreement fast vs exact:", float((out["doa"] == out2["doa"]).float().mean()),
      "counts equal frac:", float((out["counts"] == out2["counts"]).float().mean()))
