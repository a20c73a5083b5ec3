#!/bin/bash
cd /root/repo
mkdir -p gpurun_out
run() { # exe seed tracefile
  d=$(mktemp -d); ( cd $d; env SBG_SEEDFILE=/root/repo/tests/golden/$2.bin SBG_SHIM_TRACE=/root/repo/gpurun_out/$3 /root/repo/oracle/_ref/$1 -l -o 0 /root/repo/oracle/_ref/sboxes/rijndael.txt > /dev/null 2>&1; ls *.xml ); rm -rf $d; }
run sboxgates_gpu2 seed1 trace_gpu2.txt
run sboxgates_gpu seed1 trace_node.txt
wc -l gpurun_out/trace_*.txt
cmp gpurun_out/trace_gpu2.txt gpurun_out/trace_node.txt | head -2
python - <<'PY'
a=open('/root/repo/gpurun_out/trace_gpu2.txt').read().splitlines()
b=open('/root/repo/gpurun_out/trace_node.txt').read().splitlines()
for i,(x,y) in enumerate(zip(a,b)):
    if x!=y:
        print("first difference at call", i); print(" gpu2:", a[max(0,i-2):i+3]); print(" node:", b[max(0,i-2):i+3]); break
PY
head -c 100000 gpurun_out/trace_gpu2.txt > gpurun_out/t2_head.txt; head -c 100000 gpurun_out/trace_node.txt > gpurun_out/tn_head.txt; rm gpurun_out/trace_*.txt
