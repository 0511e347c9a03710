#!/bin/bash
# round-2 GPU call 15 (4 GPUs): which recurrence at 2 and 4 GPUs (4.2 M and 2.1 M rows per rank), with the evict_first matrix
# stream and the weighted ghost-CTA partition; the picpart-file parity case again
set -x
O=gpurun_out/exp15; mkdir -p $O
export OHP_P2P_TIMEOUT_S=10
timeout 300 python -m pytest tests/test_multigpu.py -m gpu -v -k "picpart_files or merged" > $O/pytest_mgpu_picpart.log 2>&1; echo "pytest rc=$?" >> $O/pytest_mgpu_picpart.log
for n in 4 2; do
  L="python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29700"
  B="bench.py --gpus $n --no-cpu --secondary none --e2e-steps 2 --steps 3 --warmup 3 --sampler none"
  OHP_BENCH_E2E_BREAKDOWN=1 timeout 300 $L $B > $O/n${n}_default.json 2>> $O/err.txt
  for cg in cgcg pipecg2; do
    OHP_CG=$cg timeout 300 $L $B --e2e-steps 1 > $O/n${n}_$cg.json 2>> $O/err.txt
  done
done
ls $O
