set -x
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv
timeout 1700 python -m pytest tests/test_gpu_parity.py tests/test_gpu_node.py -m gpu -x -q 2>&1 | tail -40
