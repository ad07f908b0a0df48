#!/usr/bin/env python
"""Split an exported ncu source page (ncu -i X.ncu-rep --page source --csv) at its BAR.SYNC instructions:
in the warp-specialised fused kernel every role's tile loop ends in its own barrier, so the pieces are the
roles.  Per piece: warp instructions, FMA-pipe cycles they need (FFMA2 = 2, other FMA-pipe ops = 1), samples
and the stall split."""
import csv
import sys
from collections import defaultdict

FMA_OPS = ("FFMA", "FMUL", "FADD", "IMAD", "HFMA2", "FSEL")  # ops ptxas sends to the fma pipe

rows = list(csv.reader(open(sys.argv[1])))
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[hi]
ix = {h: i for i, h in enumerate(hdr)}
keys = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
pieces, cur = [], None


def new():
    return {"n": 0, "s": 0, "ffma2": 0, "fma1": 0, "lds": 0, "st": defaultdict(int), "first": None, "rows": 0}


cur = new()
for i, r in enumerate(rows[hi + 1:]):
    if len(r) <= ix["stall_wait"]:
        continue
    src = r[ix["Source"]].strip()
    op = src.split()[1] if src.startswith("@") else src.split()[0]
    n, s = int(r[ix["Instructions Executed"]]), int(r[ix["# Samples"]])
    if cur["first"] is None:
        cur["first"] = i
    cur["n"] += n; cur["s"] += s; cur["rows"] += 1
    if op.startswith("FFMA2"):
        cur["ffma2"] += n
    elif op.split(".")[0] in FMA_OPS:
        cur["fma1"] += n
    if op.startswith("LDS") or op.startswith("STS"):
        cur["lds"] += n
    for k in keys:
        cur["st"][k] += int(r[ix[k]] or 0)
    if "BAR.SYNC" in src:
        pieces.append(cur); cur = new()
pieces.append(cur)
tot_n = sum(p["n"] for p in pieces); tot_s = sum(p["s"] for p in pieces)
print(f"total warp-inst {tot_n}, samples {tot_s}")
for j, p in enumerate(pieces):
    if p["n"] < tot_n / 300:
        continue
    tops = sorted(p["st"].items(), key=lambda kv: -kv[1])[:6]
    print(f"piece {j} rows {p['first']}+{p['rows']}: inst {p['n']/1e6:.1f}M ({100*p['n']/tot_n:.1f}%) ffma2 {p['ffma2']/1e6:.1f}M fma1 {p['fma1']/1e6:.1f}M "
          f"lds/sts {p['lds']/1e6:.1f}M fma-pipe-cyc {(2*p['ffma2']+p['fma1'])/1e6:.1f}M samples {p['s']} ({100*p['s']/tot_s:.1f}%): "
          + ", ".join(f"{k[6:]} {v}" for k, v in tops))
