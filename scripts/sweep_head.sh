#!/bin/bash
# Exploration aid: phase-1 kernel time vs size of the chunked head (waves of warps) and ticket batch.
for b in auto 1; do
for w in none 1 2 4 8 16 64; do
  if [ $b = auto ]; then unset SBG_BATCH; else export SBG_BATCH=$b; fi
  if [ $w = none ]; then export SBG_HEAD=0; unset SBG_HEAD_WAVES; else export SBG_HEAD=1 SBG_HEAD_WAVES=$w; fi
  echo "== batch $b head $w"; python scripts/explore_sizes.py ${@:-24 32 40 64} 2>&1 | grep -E "mask=(256| 64) " | sed -e 's/wall=[0-9.]*ms//g' -e 's/|.*| 7lut/|/' | cut -c1-120
done
done
