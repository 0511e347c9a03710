#!/bin/bash
# round-2 GPU call 23 (1 GPU): 3-D SpMV: one thread per row with fewer consumer warps (configs 6-8), 512 consumers with lane sharing (9)
set -x
O=gpurun_out/exp23; mkdir -p $O
B="python bench.py --no-cpu --secondary none --e2e-steps 0 --steps 10 --warmup 3 --sampler none --workload box3d_110"
for c in 6 7 8 9 1; do OHP_SPMV=$c timeout 300 $B > $O/3d_cfg$c.json 2>> $O/err.txt; done
ls $O
