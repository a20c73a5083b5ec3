#!/bin/bash
# Launch profile of a REAL run: `sboxgates_gpu -l -o 0 rijndael.txt` (seed1), launches 40000-44000.
cd /root/repo; mkdir -p gpurun_out
d=$(mktemp -d); cd $d
env SBG_SEEDFILE=/root/repo/tests/golden/seed1.bin ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum,launch__grid_size \
   --clock-control none -s 40000 -c 4000 --csv --log-file /root/repo/gpurun_out/r02_launches_dropin_slice.csv \
   /root/repo/oracle/_ref/sboxgates_gpu -l -o 0 /root/repo/oracle/_ref/sboxes/rijndael.txt > /dev/null 2>&1
cd /root/repo; python - <<'PY'
import csv,collections
rows=[l for l in open('/root/repo/gpurun_out/r02_launches_dropin_slice.csv') if l.startswith('"')]
r=list(csv.DictReader(rows))
agg=collections.defaultdict(lambda: collections.defaultdict(list))
for x in r:
    name=x['Kernel Name'].split('(')[0].replace('void ','').replace('sbg::','')
    agg[name][x['Metric Name']].append(float(x['Metric Value'].replace(',','')))
print("%-34s %6s %9s %8s %8s %10s %7s"%("kernel","n","total us","mean us","p90 us","mean inst","grid"))
for k,v in sorted(agg.items(), key=lambda kv:-sum(kv[1]['gpu__time_duration.sum'])):
    t=sorted(v['gpu__time_duration.sum']); n=len(t)
    print("%-34s %6d %9.0f %8.1f %8.1f %10.0f %7.0f"%(k[:34],n,sum(t)/1e3,sum(t)/n/1e3,t[int(0.9*n)]/1e3,sum(v['smsp__inst_executed.sum'])/n,sum(v['launch__grid_size'])/n))
PY
