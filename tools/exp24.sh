#!/bin/bash
# round-2 GPU call 24 (1 GPU): 3-D SpMV consumer configurations, e2e breakdown at N = 1, the new every-recurrence test, full suite
set -x
O=gpurun_out/exp24; mkdir -p $O
B="python bench.py --no-cpu --secondary none --e2e-steps 0 --steps 10 --warmup 3 --sampler none --workload box3d_110"
for c in 6 7 8 9 1; do OHP_SPMV=$c timeout 300 $B > $O/3d_cfg$c.json 2>> $O/err.txt; done
OHP_BENCH_E2E_BREAKDOWN=1 timeout 300 python bench.py --no-cpu --secondary none --e2e-steps 3 --steps 1 --warmup 1 --sampler none > $O/n1_e2e_breakdown.json 2>> $O/err.txt
OHP_BENCH_E2E_BREAKDOWN=1 timeout 300 python bench.py --no-cpu --secondary none --e2e-steps 3 --steps 3 --warmup 1 --sampler none --workload box3d_110 > $O/3d_e2e_breakdown.json 2>> $O/err.txt
timeout 1500 python -m pytest tests -m gpu -q > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $O/pytest_gpu.log
ls $O
