#!/bin/bash
# round-2 GPU call 9 (4 GPUs): multi-GPU parity, 2- and 4-GPU cases (the 8-GPU ones skip here), log kept for profiles/;
# then the N=4 bench with the single-reduction and the pipelined recurrence
set -x
O=gpurun_out/exp9; mkdir -p $O
export OHP_P2P_TIMEOUT_S=10
timeout 1500 python -m pytest tests/test_multigpu.py -m gpu -v > $O/pytest_mgpu_4gpu_box.log 2>&1; echo "pytest rc=$?" >> $O/pytest_mgpu_4gpu_box.log
L="python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29700"
B="bench.py --gpus 4 --no-cpu --secondary none --e2e-steps 2 --steps 3 --warmup 3 --sampler none"
for cg in cgcg pipecg; do
  OHP_CG=$cg timeout 300 $L $B > $O/n4_$cg.json 2>> $O/err.txt
done
ls $O
