#!/bin/bash
# 2 GPUs: the 8(f)-rank-4 slice across ranks (Chebyshev CG, natural BC, null space, Newton)
mkdir -p gpurun_out
timeout 95 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29617 tests/mgpu_next_worker.py > gpurun_out/r02_mgpu_next_rows_2gpu.log 2>&1
echo "rc=$?" >> gpurun_out/r02_mgpu_next_rows_2gpu.log
grep -v "OMP_NUM_THREADS\|\*\*\*\*" gpurun_out/r02_mgpu_next_rows_2gpu.log | tail -25
