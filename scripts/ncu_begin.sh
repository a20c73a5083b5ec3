#!/bin/bash
cd /root/repo; mkdir -p gpurun_out
d=$(mktemp -d); cd $d
env SBG_SEEDFILE=/root/repo/tests/golden/seed1.bin ncu --set full --clock-control none --import-source on -k regex:k_begin -s 9000 -c 3 \
   -o /root/repo/gpurun_out/r02_begin /root/repo/oracle/_ref/sboxgates_gpu -l -o 0 /root/repo/oracle/_ref/sboxes/rijndael.txt > /dev/null 2>&1
ls -la /root/repo/gpurun_out/r02_begin.ncu-rep
