#!/bin/bash
# Exploration aid: phase-1 kernel time of the position-major kernel with 4- vs 5-gate prefixes.
for p in 4 5; do
  export SBG_PM_PREFIX=$p
  echo "== prefix $p"; python scripts/explore_sizes.py 24 32 40 48 64 96 2>&1 | grep -E "mask=(256| 64) " | sed -e 's/wall=[0-9.]*ms//g' -e 's/5lut.*| 7lut/7lut/' | cut -c1-150
done
