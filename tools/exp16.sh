#!/bin/bash
# round-2 GPU call 16 (1 GPU): cell-parallel (patch) assembly: parity suite, then assembly kernel times against the row kernels
set -x
O=gpurun_out/exp16; mkdir -p $O
timeout 1200 python -m pytest tests -m gpu -q -x > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $O/pytest_gpu.log
B="python bench.py --no-cpu --secondary none --e2e-steps 2 --steps 3 --warmup 3 --sampler none"
timeout 300 $B > $O/full_patch.json 2>> $O/err.txt
OHP_ASM=rows timeout 300 $B > $O/full_rows.json 2>> $O/err.txt
timeout 300 $B --workload box3d_110 --steps 10 > $O/3d_patch.json 2>> $O/err.txt
OHP_ASM=rows timeout 300 $B --workload box3d_110 --steps 10 > $O/3d_rows.json 2>> $O/err.txt
ls $O
