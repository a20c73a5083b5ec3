#!/bin/bash
cd /root/repo; mkdir -p gpurun_out
for mode in 0 1 32; do
  SBG_DECOMP_FILTER=$mode ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum,sm__warps_active.avg.pct_of_peak_sustained_active,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__cycles_active.avg,launch__grid_size \
    --clock-control none -k regex:k_decomp7 --csv --log-file gpurun_out/decomp_$mode.csv python scripts/one_decomp.py > /dev/null 2>&1
  python - <<PY
import csv
rows=[l for l in open('/root/repo/gpurun_out/decomp_$mode.csv') if l.startswith('"')]
r=list(csv.DictReader(rows))
by={}
for x in r: by.setdefault(x['ID'],{})[x['Metric Name']]=x['Metric Value']
for k,v in by.items(): print('filter=$mode', k, v)
PY
done
