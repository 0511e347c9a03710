#!/bin/bash
for v in nopdl pdl; do
  if [ $v = nopdl ]; then export OHP_NOPDL=1; else unset OHP_NOPDL; fi
  OHP_CG=classic timeout 200 python bench.py --steps 2 --warmup 2 --no-cpu --e2e-steps 1 --n $1 2>&1 | tail -1 | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('$v', 'n=$1 step_ms %.1f'%d['ms_per_step'], 'cg_iter_ms %.4f'%d['kernels']['cg_iteration']['ms'])"
done
