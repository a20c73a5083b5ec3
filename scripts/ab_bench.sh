#!/bin/bash
# A/B: bench.py (headline only) under several builds of the library (SBG_LIB=...).
for spec in "$@"; do
  lib=${spec%%:*}; envs=${spec#*:}; [ "$envs" = "$spec" ] && envs=""
  echo "== $lib $envs"
  env SBG_LIB=$lib $envs python bench.py --no-extras --no-cpu-baseline --steps 10 --warmup 3 2>/dev/null | python -c "
import json,sys
l=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('ms/step %.3f  e2e %.3f  isolated %s  value %.3e' % (l['ms_per_step'], l['e2e']['ms_per_step'], {k: round(v,3) for k,v in l['kernel_ms_per_step_isolated'].items()}, l['value']))
print('   by mask depth', {k: [round(x,3) for x in v] for k,v in l['kernel_ms_per_step_by_mask_depth'].items()})"
done
