#!/bin/bash
# scratch driver for one gpurun visit (edited per call)
mkdir -p gpurun_out
timeout 900 ncu --metrics sm__cycles_elapsed.max,smsp__inst_executed.sum,gpu__time_duration.sum,sm__issue_active.avg.pct_of_peak_sustained_elapsed,l1tex__throughput.avg.pct_of_peak_sustained_elapsed --clock-control none -k regex:conv0_tc_pool -c 64 --csv --log-file gpurun_out/conv0_variants.csv python tools/ab_conv0.py > gpurun_out/c35_ncu.log 2>&1; echo "ncu rc=$?"; tail -3 gpurun_out/c35_ncu.log
python - <<'PY'
import csv
rows=[r for r in csv.reader(open('gpurun_out/conv0_variants.csv')) if len(r)>10]
h=rows[0]; ik=h.index('Kernel Name'); im=h.index('Metric Name'); iv=h.index('Metric Value'); iid=h.index('ID')
d={}
for r in rows[1:]:
    d.setdefault(int(r[iid]),{'k':r[ik][:48]})[r[im]]=r[iv]
for i in sorted(d):
    if i in (1,2,5,10,20,25,30,40,45,50,60,63): print(i, d[i])
PY
