#!/bin/bash
# round-2 GPU call 18 (2 GPUs): exchange-heap cache (reuse across meshes in the e2e loop; drop at context teardown)
set -x
O=gpurun_out/exp18; mkdir -p $O
export OHP_P2P_TIMEOUT_S=10
timeout 600 python -m pytest tests/test_multigpu.py -m gpu -v -k "two_gpus and (box2d or 24k_4p) or cxx_picpart_two or merged" > $O/pytest_mgpu2.log 2>&1; echo "pytest rc=$?" >> $O/pytest_mgpu2.log
L="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29700"
B="bench.py --gpus 2 --no-cpu --secondary none --e2e-steps 3 --steps 2 --warmup 2 --sampler none"
OHP_BENCH_E2E_BREAKDOWN=1 timeout 300 $L $B > $O/n2_cache.json 2>> $O/err.txt
OHP_P2P_NOCACHE=1 OHP_BENCH_E2E_BREAKDOWN=1 timeout 300 $L $B > $O/n2_nocache.json 2>> $O/err.txt
timeout 300 $L $B --workload box3d_110 --steps 10 > $O/n2_3d.json 2>> $O/err.txt
ls $O
