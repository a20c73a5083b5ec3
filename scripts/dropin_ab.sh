#!/bin/bash
# A/B one shim environment variable on a real seeded run of the drop-in CLI (run on the GPU box):
#   scripts/dropin_ab.sh VAR value_a value_b
var=${1:-SBG_HEAD}; a=${2:-0}; b=${3:-1}
R=$GRAFT_REPO_ROOT
for rep in 1 2; do
for v in $a $b; do
  rm -rf /tmp/ab; mkdir -p /tmp/ab; cd /tmp/ab
  s=$(date +%s%N)
  env $var=$v SBG_SEEDFILE=$R/tests/golden/seed1.bin SBG_SHIM_STATS=1 $R/oracle/_ref/sboxgates_gpu -l -o 0 $R/oracle/_ref/sboxes/rijndael.txt > out.txt 2>&1
  e=$(date +%s%N)
  grep -i "search_\|kernel\|calls" out.txt | tail -n 6
  echo "$var=$v wall_ms $(( (e - s) / 1000000 )) $(ls *.xml | tail -n 1)"
done
done
