#!/bin/bash
# round-2 GPU call 12 (8 GPUs): 8-GPU parity cases (log kept for profiles/), then config 4 at N=8 with the three
# recurrences (single-reduction two-launch, pipelined two-launch, pipelined one-launch) and a per-rank device timeline
set -x
O=gpurun_out/exp12; mkdir -p $O
export OHP_P2P_TIMEOUT_S=10
timeout 600 python -m pytest tests/test_multigpu.py -m gpu -v -k "eight" > $O/pytest_mgpu_8gpu_box.log 2>&1; echo "pytest rc=$?" >> $O/pytest_mgpu_8gpu_box.log
L="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29700"
B="bench.py --gpus 8 --no-cpu --secondary none --e2e-steps 2 --steps 3 --warmup 3 --sampler none"
for cg in pipecg cgcg pipecg2; do
  OHP_CG=$cg timeout 200 $L $B > $O/n8_$cg.json 2>> $O/err.txt
done
OHP_CG=pipecg OHP_TRACE_CG=$O/trace_n8_pipecg timeout 200 $L $B --e2e-steps 0 --steps 1 --warmup 1 > $O/trace_run.json 2>> $O/err.txt
ls $O
