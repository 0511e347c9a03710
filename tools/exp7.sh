#!/bin/bash
# round-2 GPU call 7 (2 GPUs): multi-GPU parity (2-GPU cases, now including pipecg + Jacobi solves) and the N=2 bench
# with the single-reduction and the pipelined recurrence
set -x
O=gpurun_out/exp7; mkdir -p $O
export OHP_P2P_TIMEOUT_S=10
timeout 900 python -m pytest tests/test_multigpu.py -m gpu -q -x > $O/pytest_mgpu2.log 2>&1; echo "pytest rc=$?" >> $O/pytest_mgpu2.log
L="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29700"
B="bench.py --gpus 2 --no-cpu --secondary none --e2e-steps 1 --steps 3 --warmup 3 --sampler none"
for cg in cgcg pipecg; do
  OHP_CG=$cg timeout 300 $L $B > $O/n2_$cg.json 2>> $O/err.txt
done
OHP_CG=pipecg OHP_TRACE_CG=$O/trace_n2_pipecg timeout 300 $L $B --steps 1 --warmup 1 > $O/trace_run.json 2>> $O/err.txt
ls $O
