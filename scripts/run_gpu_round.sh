#!/bin/bash
# One GPU-box session: the whole GPU test-suite, then the bench with all its records.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv
( timeout 2400 python -m pytest tests -m gpu -x -q 2>&1 | tail -25 ) > gpurun_out/tests.log 2>&1
tail -8 gpurun_out/tests.log
( time timeout 1200 python bench.py ) > gpurun_out/bench.log 2> gpurun_out/bench.err
tail -c 3000 gpurun_out/bench.log; tail -5 gpurun_out/bench.err
