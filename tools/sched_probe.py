#!/usr/bin/env python
"""Scheduler probe scenarios (runs on the GPU box): how FFMA2 streams, dependent chains and ALU streams
share one SM sub-partition.  Prints cycles per instruction of every active warp."""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from haghighatshoarmuir2024_b200 import _native as N
lib = N.lib()
NAMES = {0: "-", 1: "ffma2", 2: "chain", 3: "alu", 4: "ffma", 5: "chainmix", 6: "ffma2+alu", 7: "dchain", 8: "dfma", 9: "fir2", 10: "fir1", 11: "fir2p", 12: "fir2u", 13: "fir2uw", 14: "fir2w"}

def run(title, placement, iters=4000):
    roles = (ctypes.c_int32 * 16)(*([0] * 16))
    for w, r in placement.items():
        roles[w] = r
    out = (ctypes.c_uint64 * 32)()
    N.check(lib.micloc_sched_probe(0, roles, iters, out))
    parts = []
    for w in range(16):
        if roles[w]:
            parts.append(f"w{w} {NAMES[roles[w]]}: {out[16 + w] / max(out[w], 1):.3f}")
    print(f"{title:46s} | " + "  ".join(parts) + "  inst/cyc")

run("ffma2 alone", {0: 1})
run("ffma2 x2 same SMSP", {0: 1, 4: 1})
run("ffma2 x3 same SMSP", {0: 1, 4: 1, 8: 1})
run("chain alone", {0: 2})
run("alu alone", {0: 3})
run("ffma alone", {0: 4})
run("ffma x2 same SMSP", {0: 4, 4: 4})
run("ffma2 + alu (alu higher wid)", {0: 1, 4: 3})
run("alu + ffma2 (ffma2 higher wid)", {0: 3, 4: 1})
run("ffma2 + ffma", {0: 1, 4: 4})
run("ffma2 + chain (chain higher wid)", {0: 1, 4: 2})
run("chain + ffma2 (chain lower wid)", {0: 2, 4: 1})
run("ffma2 x2 + chain highest", {0: 1, 4: 1, 8: 2})
run("chain lowest + ffma2 x2", {0: 2, 4: 1, 8: 1})
run("ffma2 x2 + chain + chainmix (high)", {0: 1, 4: 1, 8: 2, 12: 5})
run("chain + chainmix (low) + ffma2 x2", {0: 2, 4: 5, 8: 1, 12: 1})
run("ffma2+alu x2 + chain + chainmix (high)", {0: 6, 4: 6, 8: 2, 12: 5})
run("ffma2 x2 + alu highest", {0: 1, 4: 1, 8: 3})
run("ffma2 x1 + chain x2", {0: 1, 4: 2, 8: 2})
run("ffma x2 + chain highest", {0: 4, 4: 4, 8: 2})
run("chain lowest + ffma x2", {0: 2, 4: 4, 8: 4})
run("ffma x2 + chain + chainmix (high)", {0: 4, 4: 4, 8: 2, 12: 5})
run("ffma + alu", {0: 4, 4: 3})
run("ffma2+alu x1 + chain + chainmix", {0: 6, 8: 2, 12: 5})
run("ffma2 + chain + chainmix", {0: 1, 8: 2, 12: 5})
run("ffma + chain + chainmix", {0: 4, 8: 2, 12: 5})
run("dchain alone", {0: 7})
run("dfma alone", {0: 8})
run("dfma x2", {0: 8, 4: 8})
run("ffma2 + dfma", {0: 1, 4: 8})
run("ffma2 x2 + dfma", {0: 1, 4: 1, 8: 8})
run("ffma2 x2 + dchain", {0: 1, 4: 1, 8: 7})
run("ffma2 x2 + dchain x2", {0: 1, 4: 1, 8: 7, 12: 7})
run("ffma x2 + dchain x2", {0: 4, 4: 4, 8: 7, 12: 7})
run("ffma2 x2 + dchain + chainmix", {0: 1, 4: 1, 8: 7, 12: 5})
run("FIR-pattern ffma2 (scalar tap) x1", {0: 9})
run("FIR-pattern ffma2 (scalar tap) x2", {0: 9, 4: 9})
run("FIR-pattern ffma2 (scalar tap) x3", {0: 9, 4: 9, 8: 9})
run("FIR-pattern ffma2 (pair tap) x1", {0: 11})
run("FIR-pattern ffma2 (pair tap) x2", {0: 11, 4: 11})
run("FIR-pattern ffma2 (pair tap) x3", {0: 11, 4: 11, 8: 11})
run("FIR-pattern ffma x1", {0: 10})
run("FIR-pattern ffma x2", {0: 10, 4: 10})
run("FIR-pattern ffma x3", {0: 10, 4: 10, 8: 10})
run("FIR-pattern ffma2 x3, all 4 SMSPs", {w: 9 for w in range(12)})
run("FIR ffma2, uniform taps, tap-major x1", {0: 12})
run("FIR ffma2, uniform taps, tap-major x2", {0: 12, 4: 12})
run("FIR ffma2, uniform taps, window-major x1", {0: 13})
run("FIR ffma2, uniform taps, window-major x2", {0: 13, 4: 13})
run("FIR ffma2, register taps, window-major x1", {0: 14})
run("FIR ffma2, register taps, window-major x2", {0: 14, 4: 14})
