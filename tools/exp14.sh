#!/bin/bash
# round-2 GPU call 14 (8 GPUs): 8-GPU parity cases + the picpart-file case again (weighted ghost-CTA partition, pipecg on a
# rank without rows), then config 4 at N=8: library default (expected: two-launch pipelined CG), uniform partition A/B, one-launch form
set -x
O=gpurun_out/exp14; mkdir -p $O
export OHP_P2P_TIMEOUT_S=10
timeout 600 python -m pytest tests/test_multigpu.py -m gpu -v -k "eight or picpart_files" > $O/pytest_mgpu_8gpu_box.log 2>&1; echo "pytest rc=$?" >> $O/pytest_mgpu_8gpu_box.log
L="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29700"
B="bench.py --gpus 8 --no-cpu --secondary none --e2e-steps 3 --steps 3 --warmup 3 --sampler none"
OHP_BENCH_E2E_BREAKDOWN=1 timeout 200 $L $B > $O/n8_default.json 2>> $O/err.txt
OHP_GHOST_CTA_SHARE=100 timeout 200 $L $B --e2e-steps 1 > $O/n8_uniform.json 2>> $O/err.txt
OHP_CG=pipecg1 timeout 200 $L $B --e2e-steps 1 > $O/n8_pipecg1.json 2>> $O/err.txt
ls $O
