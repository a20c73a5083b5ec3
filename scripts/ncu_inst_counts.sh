#!/bin/bash
# Warp-instruction counts of the dominant kernel (k_filter7_pm) on the very states bench.py times:
# ncu profiles every k_filter7_pm launch of a bench run (same seeds, same order), and
# scripts/parse_inst_counts.py turns the CSV into profiles/r02_filter_inst_counts.json, which
# bench.py divides by its own CUDA-event kernel times for the ALU roofline.
#   usage (GPU box): bash scripts/ncu_inst_counts.sh [steps] [warmup] [gates] [batch]
set -e
STEPS=${1:-10}; WARMUP=${2:-3}; GATES=${3:-40}; BATCH=${4:-8}
mkdir -p gpurun_out
ncu --metrics smsp__inst_executed.sum,gpu__time_duration.sum,smsp__inst_executed_pipe_alu.sum,launch__registers_per_thread \
    --clock-control none -k regex:k_filter7_pm --csv --log-file gpurun_out/filter_inst.csv \
    python bench.py --steps $STEPS --warmup $WARMUP --gates $GATES --batch $BATCH --no-extras --no-cpu-baseline \
    > gpurun_out/filter_inst_bench.log 2>&1
python scripts/parse_inst_counts.py gpurun_out/filter_inst.csv $STEPS $WARMUP $GATES $BATCH \
    > gpurun_out/r02_filter_inst_counts.json
cp gpurun_out/r02_filter_inst_counts.json profiles/r02_filter_inst_counts.json
tail -c 600 gpurun_out/r02_filter_inst_counts.json
