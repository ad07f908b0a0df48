#!/bin/bash
# per-kernel times of the Xylo chain (fast / LIF only / exact) at B clips: ncu launch list + the plain bench
B=${1:-444}; TAG=${2:-x}
mkdir -p gpurun_out
python tools/xylo_bench.py $B > gpurun_out/xylo_bench_$TAG.log 2>&1; cat gpurun_out/xylo_bench_$TAG.log
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/xylo_launches_$TAG.csv python tools/xylo_bench.py $B > gpurun_out/xylo_ncu_$TAG.log 2>&1
python - <<PY
import csv, collections
rows = list(csv.reader(l for l in open("gpurun_out/xylo_launches_$TAG.csv") if l.startswith('"')))
hdr = rows[0]; ki = hdr.index("Kernel Name"); vi = hdr.index("Metric Value"); ui = hdr.index("Metric Unit")
agg = collections.OrderedDict()
for r in rows[1:]:
    name = r[ki].split("(")[0][:60]
    v = float(r[vi].replace(",", "")); u = r[ui]
    v = v / 1e6 if u in ("ns", "nsecond") else (v / 1e3 if u in ("us", "usecond") else v)
    a = agg.setdefault(name, [0, 0.0]); a[0] += 1; a[1] += v
for k, (n, ms) in agg.items(): print(f"{ms:10.3f} ms  {n:4d} x  {k}")
PY
