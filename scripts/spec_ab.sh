#!/bin/bash
cd /root/repo
for sp in 1 0 2 1; do
  EXTRA_ENV="SBG_SPECULATE=$sp" bash scripts/dropin_time.sh sboxgates_gpu 2>&1 | grep -E "^==|node calls|waiting by" | sed "s/^/[spec=$sp] /" | cut -c1-230
done
