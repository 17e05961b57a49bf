#!/bin/bash
# true per-kernel durations of the ABI request shapes (ncu, no clock control); $1 = tag
mkdir -p gpurun_out
for cfg in "16000 nt 200" "20000 aa 1287"; do
  set -- $cfg
  timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/kt_$2.csv python profiles/kernel_times.py $1 $2 $3 > gpurun_out/kt_$2.log 2>&1
  python - <<PY
import csv
rows=[r for r in csv.reader(open('gpurun_out/kt_$2.csv')) if len(r)>10]
hdr=rows[0]; ki=hdr.index('Kernel Name'); vi=hdr.index('Metric Value'); gi=hdr.index('Grid Size')
out=[]
for r in rows[1:]:
    nm=r[ki].split('<')[0].split('(')[0]
    out.append((nm, r[gi], float(r[vi].replace(',',''))))
# only the tail after the setup joins: print the last 60 launches
for nm,g,v in out[-52:]:
    print('$2 %-24s grid %-14s %10.2f us' % (nm, g, v/1000 if v>1e3 else v))
PY
done
