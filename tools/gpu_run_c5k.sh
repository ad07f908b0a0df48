ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_c5m.csv python tools/c5_probe.py 256 48000 > /dev/null 2>&1
python - <<PY
import csv
rows=list(csv.reader(open('gpurun_out/launches_c5m.csv')))
hi=[i for i,r in enumerate(rows) if r and r[0]=='ID'][0]; hdr=rows[hi]; ki=hdr.index('Kernel Name'); vi=hdr.index('Metric Value')
seq=[(r[ki][:70], float(r[vi].replace(',',''))/1e6) for r in rows[hi+1:] if len(r)>vi]
n=len(seq)//3
for k,t in seq[-n:]: print(f"{t:9.3f} ms  {k}")
PY
