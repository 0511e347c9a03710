#!/bin/bash
# classic vs persistent CG at a given box size on one GPU
for c in classic persistent; do OHP_CG=$c timeout 200 python bench.py --steps 2 --warmup 2 --no-cpu --e2e-steps 1 --n $1 2>&1 | tail -1 | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('$c', 'n=$1 step_ms %.1f'%d['ms_per_step'], 'cg_iter_ms %.4f'%d['kernels']['cg_iteration']['ms'], 'its %d'%d['kernels']['cg_iteration']['iterations_per_step'], d['kernels']['phase_ms_per_step'])"; done
