#!/usr/bin/env python
"""Join an ncu SASS-level source page with nvdisasm line info: per CUDA source line,
warp instructions executed and stall samples.  Runs here (no GPU needed).

    python tools/ncu_lines.py gpurun_out/prof.ncu-rep haghighatshoarmuir2024_b200/csrc/micloc_fused.o k_fusedIf [top]

Set NCU_RANGES="name:lo-hi,name:lo-hi" (line ranges of micloc_fused.cu) to get, per range, the
instruction count and the stall-reason split of its samples (e.g. one range per warp role).
"""
import csv
import io
import os
import re
import subprocess
import sys
import tempfile
from collections import defaultdict


def main():
    rep, obj, kern = sys.argv[1], sys.argv[2], sys.argv[3]
    top = int(sys.argv[4]) if len(sys.argv) > 4 else 45
    txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
    hdr_i = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
    hdr = rows[hdr_i]
    ci, cs, cx = hdr.index("Instructions Executed"), hdr.index("# Samples"), hdr.index("Source")
    cth = hdr.index("Thread Instructions Executed")
    inst = [(r[cx].strip(), int(r[ci]), int(r[cs]), int(r[cth])) for r in rows[hdr_i + 1:] if len(r) > cth]
    stall_cols = [(h, i) for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
    stalls = [[int(r[i] or 0) for _, i in stall_cols] for r in rows[hdr_i + 1:] if len(r) > cth]
    with tempfile.TemporaryDirectory() as td:
        subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(obj)], cwd=td, capture_output=True)
        cub = [f for f in os.listdir(td) if f.endswith(".cubin")][0]
        dis = subprocess.run(["nvdisasm", "--print-line-info", "-c", os.path.join(td, cub)], capture_output=True,
                             text=True).stdout.splitlines()
    # walk the function's text
    lines = []
    infn = False
    cur = ("?", 0)
    for ln in dis:
        if ln.startswith(".text."):
            infn = kern in ln
            continue
        if not infn:
            continue
        m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
        if m:
            cur = (os.path.basename(m.group(1)), int(m.group(2)))
            continue
        m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(.*?);", ln)
        if m:
            lines.append((cur, m.group(1).strip()))
    if len(lines) != len(inst):
        print(f"warning: {len(lines)} disassembled vs {len(inst)} profiled instructions", file=sys.stderr)
    agg = defaultdict(lambda: [0, 0, 0])
    tot_i = tot_s = 0
    for (loc, _), (_, n, s, th) in zip(lines, inst):
        agg[loc][0] += n; agg[loc][1] += s; agg[loc][2] += th
        tot_i += n; tot_s += s
    print(f"total warp instructions {tot_i}, samples {tot_s}")
    # split the SASS at its BAR.SYNC instructions: in a warp-specialised kernel each role's loop
    # body ends in its own barrier, so the pieces are (roughly) the roles
    seg, segs = [0, 0, [0] * len(stall_cols), defaultdict(int)], []
    for (loc, txt_i), (_, n, sm, _), sv in zip(lines, inst, stalls):
        seg[0] += n; seg[1] += sm
        seg[2] = [a + b for a, b in zip(seg[2], sv)]
        seg[3][loc] += n
        if txt_i.startswith("BAR.SYNC") or "BAR.SYNC" in txt_i:
            segs.append(seg)
            seg = [0, 0, [0] * len(stall_cols), defaultdict(int)]
    segs.append(seg)
    for i, (ni, ns, st, locs) in enumerate(segs):
        if ni < tot_i / 200:
            continue
        tops = sorted(zip([h for h, _ in stall_cols], st), key=lambda kv: -kv[1])[:6]
        main = max(locs.items(), key=lambda kv: kv[1])[0]
        print(f"  piece {i} (mostly {main[0]}:{main[1]}): warp-inst {ni} ({100 * ni / tot_i:.1f}%), samples {ns} "
              f"({100 * ns / max(tot_s, 1):.1f}%): " + ", ".join(f"{h[6:]} {v}" for h, v in tops))
    ranges = os.environ.get("NCU_RANGES")
    if ranges:
        main_file = os.environ.get("NCU_FILE", "micloc_fused.cu")
        for spec in ranges.split(","):
            name, lohi = spec.split(":")
            lo, hi = map(int, lohi.split("-"))
            ni = ns = 0
            st = [0] * len(stall_cols)
            for ((f, ln), _), (_, n, sm, _), sv in zip(lines, inst, stalls):
                if f == main_file and lo <= ln <= hi:
                    ni += n; ns += sm
                    st = [a + b for a, b in zip(st, sv)]
            tops = sorted(zip([h for h, _ in stall_cols], st), key=lambda kv: -kv[1])[:6]
            print(f"  {name:8s} lines {lo}-{hi}: warp-inst {ni} ({100 * ni / tot_i:.1f}%), samples {ns} "
                  f"({100 * ns / max(tot_s, 1):.1f}%): " + ", ".join(f"{h[6:]} {v}" for h, v in tops))
    print(f"{'file:line':34s} {'warp-inst':>12s} {'%':>6s} {'samples':>9s} {'%':>6s} {'lanes':>6s}")
    for loc, (n, s, th) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
        print(f"{loc[0] + ':' + str(loc[1]):34s} {n:12d} {100 * n / tot_i:6.2f} {s:9d} {100 * s / max(tot_s, 1):6.2f} "
              f"{th / max(n, 1):6.1f}")


if __name__ == "__main__":
    main()
