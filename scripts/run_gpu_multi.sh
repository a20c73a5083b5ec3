#!/bin/bash
# Multi-GPU session (gpurun --gpus N): the sharded drop-in test, then bench.py at N ranks.
N=${1:-2}
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv,noheader | head -8
( timeout 900 python -m pytest tests/test_dropin_gpu.py -m gpu -x -q -k sharded 2>&1 | tail -8 ) > gpurun_out/tests_multi_$N.log 2>&1
tail -4 gpurun_out/tests_multi_$N.log
( time timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --steps 10 --warmup 3 ) > gpurun_out/bench_$N.log 2> gpurun_out/bench_$N.err
tail -c 2500 gpurun_out/bench_$N.log; tail -5 gpurun_out/bench_$N.err
