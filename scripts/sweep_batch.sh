#!/bin/bash
# Exploration aid: filter/search5 kernel time vs ticket batch size.
for b in 1 2 4 8 16 auto; do
  if [ $b = auto ]; then unset SBG_BATCH; else export SBG_BATCH=$b; fi
  echo "== batch $b"; python scripts/explore_sizes.py 24 32 40 64 96 2>&1 | grep -E "mask=(256| 64) " | sed -e 's/wall=[0-9.]*ms//g' | cut -c1-200
done
