#!/bin/bash
# Profiles of the bench workload on one GPU: instruction counts of the dominant kernel on the bench
# states, the launch list of a bench run, one full capture of the top kernel.
mkdir -p gpurun_out
bash scripts/ncu_inst_counts.sh 10 3 40 8
ncu --metrics gpu__time_duration.sum --clock-control none -s 150 -c 400 --csv \
    --log-file gpurun_out/r02_launches_bench_n40.csv python bench.py --steps 4 --warmup 3 --no-extras --no-cpu-baseline > gpurun_out/launches_bench.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_filter7_pm -s 24 -c 4 \
    -o gpurun_out/r02_filter7 python bench.py --steps 2 --warmup 3 --no-extras --no-cpu-baseline > gpurun_out/full_bench.log 2>&1
ls -la gpurun_out | tail -12
( time python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-extras ) 2>&1 | tail -c 1500
