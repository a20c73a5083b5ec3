#!/bin/bash
# Wall-clock to graph of the drop-in CLIs on BASELINE.json configs[1] (rijndael -l -o 0), with the
# shim's own accounting.  usage: dropin_time.sh [exe ...]   (default: sboxgates_gpu sboxgates_gpu2)
cd /root/repo
for exe in ${@:-sboxgates_gpu sboxgates_gpu2}; do
  for seed in seed1 seed2; do
    d=$(mktemp -d)
    ( cd $d; t0=$(date +%s.%N); env SBG_SEEDFILE=/root/repo/tests/golden/$seed.bin SBG_SHIM_STATS=1 $EXTRA_ENV \
        /root/repo/oracle/_ref/$exe -l -o 0 /root/repo/oracle/_ref/sboxes/rijndael.txt > /dev/null 2> err.txt
      t1=$(date +%s.%N); echo "== $exe $seed: $(ls *.xml | tr '\n' ' ') $(python3 -c "print(round($t1 - $t0, 3))") s wall"; grep "^\[sbg\]" err.txt | cut -c1-420 )
    rm -rf $d
  done
done
