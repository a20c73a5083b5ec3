#!/bin/bash
cd /root/repo
( timeout 1500 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "phase2 or reference_cases or paths_agree or maximum" 2>&1 | tail -4 )
L=sboxgates_b200/libsboxgates_b200.so
bash scripts/ab_bench.sh $L:SBG_DECOMP_FILTER=0 $L:SBG_DECOMP_FILTER=1
