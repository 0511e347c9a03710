#!/bin/bash
# A/B the SpMV kernel configurations on the full-size workload (run on the GPU box): tools/sweep_spmv.sh "0 4 5"
for cfg in $1; do
  OHP_SPMV=$cfg timeout 300 python bench.py --steps 1 --warmup 1 --no-cpu --e2e-steps 1 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('cfg $cfg', 'step_ms %.1f'%d['ms_per_step'], 'spmv_ms %.4f frac %.3f'%(d['roofline']['ms_per_launch'], d['roofline']['frac']), 'cg_iter_ms %.4f'%d['kernels']['cg_iteration']['ms'])"
done
