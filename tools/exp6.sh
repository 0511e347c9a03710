#!/bin/bash
# round-2 GPU call 6 (1 GPU): pipelined CG with residual replacement -- parity tests, then speed / true-residual check at
# the 4-, 2- and 1-GPU shares of config 4 and on the 3-D box
set -x
O=gpurun_out/exp6; mkdir -p $O
timeout 1200 python -m pytest tests -m gpu -q > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $O/pytest_gpu.log
B="python bench.py --no-cpu --secondary none --e2e-steps 1 --steps 3 --warmup 3 --sampler none"
for n in 1024 1448 2048; do
  OHP_CG=pipecg timeout 300 $B --n $n > $O/n${n}_pipecg.json 2>> $O/err.txt
done
OHP_CG=pipecg OHP_PIPE_RR=100 timeout 300 $B --n 1024 > $O/n1024_pipecg_rr100.json 2>> $O/err.txt
OHP_CG=pipecg OHP_PIPE_RR=400 timeout 300 $B --n 2048 > $O/n2048_pipecg_rr400.json 2>> $O/err.txt
OHP_CG=pipecg timeout 300 $B > $O/full_pipecg.json 2>> $O/err.txt
OHP_CG=pipecg OHP_PIPE_RR=500 timeout 300 $B > $O/full_pipecg_rr500.json 2>> $O/err.txt
OHP_CG=pipecg timeout 300 $B --workload box3d_110 --steps 10 > $O/3d_pipecg.json 2>> $O/err.txt
ls $O
