#!/bin/bash
cd /root/repo; mkdir -p gpurun_out
for bal in 1 0; do
SBG_BALANCED5=$bal ncu --set full --clock-control none --import-source on -k regex:k_sweep -o gpurun_out/r02_sweep5_bal$bal -f python scripts/one_search5.py > gpurun_out/sweep5_$bal.log 2>&1
ncu -i gpurun_out/r02_sweep5_bal$bal.ncu-rep --page raw --csv > gpurun_out/sweep5_raw_$bal.csv 2>/dev/null
python - <<PY
import csv
rows=list(csv.reader(open('/root/repo/gpurun_out/sweep5_raw_$bal.csv')))
hdr=rows[0]; idx={h:i for i,h in enumerate(hdr)}
keys=['Kernel Name','gpu__time_duration.sum','launch__grid_size','launch__registers_per_thread','smsp__inst_executed.sum','sm__warps_active.avg.pct_of_peak_sustained_active','smsp__issue_active.avg.pct_of_peak_sustained_active','sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active','sm__cycles_active.avg','gpc__cycles_elapsed.max']
keys+=[h for h in hdr if 'issue_stalled' in h and 'per_issue_active' in h]
for r in rows[2:]:
    print('== bal=$bal')
    for k in keys:
        v=r[idx[k]]
        try:
            if float(v.replace(',',''))<0.05: continue
        except: pass
        print('  ',k.replace('smsp__average_warps_issue_stalled_','stall:').replace('_per_issue_active.ratio',''),'=',v[:50])
PY
done
