#!/usr/bin/env python
"""One exact Xylo pass on B clips (for ncu captures of k_xylo_front_f64 / k_xylo_lif)."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import helpers as H
B = int(sys.argv[1]) if len(sys.argv) > 1 else 296
g = H.load("xylo_c3_bipolar")
eng = H.xylo_engine(g)
x = torch.from_numpy(H.xylo_synth_clips(g, 8, 48_000, seed=1, int16=True)).cuda()
x = x.repeat((B + 7) // 8, 1, 1)[:B].contiguous()
for _ in range(2):
    out = eng.run(x, exact=True)
torch.cuda.synchronize()
print("ok", int(out["flags"].sum()))
