#!/bin/bash
# One `ncu --set full` capture (with source correlation) of the two sweeping kernels on a bench step.
mkdir -p gpurun_out
TAG=${1:-r02c}
ncu --set full --clock-control none --import-source on -k regex:k_filter7_pm -s 24 -c 4 \
    -o gpurun_out/${TAG}_filter7 -f python bench.py --steps 2 --warmup 3 --no-extras --no-cpu-baseline > gpurun_out/full_bench.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_sweep -s 24 -c 2 \
    -o gpurun_out/${TAG}_sweep5 -f python bench.py --steps 2 --warmup 3 --no-extras --no-cpu-baseline > gpurun_out/full_bench5.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_decomp7 -s 27 -c 1 \
    -o gpurun_out/${TAG}_decomp7 -f python bench.py --steps 2 --warmup 3 --no-extras --no-cpu-baseline > gpurun_out/full_bench7.log 2>&1
ls -la gpurun_out/${TAG}_*
