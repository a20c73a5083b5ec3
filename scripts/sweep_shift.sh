#!/bin/bash
# Exploration aid: phase-1 kernel time with aligned two-word windows (SBG_SHIFT=0) and shifted one-word windows (1).
for v in 0 1; do
  echo "== SBG_SHIFT=$v"; SBG_SHIFT=$v python scripts/explore_sizes.py ${@:-24 32 40 48 56 63} 2>&1 | grep -E "mask=(256|128| 64| 32) " | sed -e 's/wall=[0-9.]*ms//g' -e 's/|.*| 7lut/|/' | cut -c1-100
done
