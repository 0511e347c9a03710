#!/bin/bash
# round-2 GPU call 11 (1 GPU): fused pipelined CG with consumer groups on different tiles + operand prefetch + register-resident scalars
set -x
O=gpurun_out/exp11; mkdir -p $O
timeout 1200 python -m pytest tests/test_gpu_parity.py tests/test_facade.py -m gpu -q > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $O/pytest_gpu.log
B="python bench.py --no-cpu --secondary none --e2e-steps 0 --steps 3 --warmup 3 --sampler none"
for n in 1024 1448 2048; do
  for g in 1 2 3; do
    OHP_CG=pipecg OHP_PF_GROUPS=$g timeout 300 $B --n $n > $O/n${n}_pipecg_g$g.json 2>> $O/err.txt
  done
done
OHP_CG=pipecg timeout 300 $B > $O/full_pipecg.json 2>> $O/err.txt
OHP_CG=pipecg OHP_PF_GROUPS=3 timeout 300 $B > $O/full_pipecg_g3.json 2>> $O/err.txt
OHP_CG=pipecg OHP_TRACE_CG=$O/trace_n1024_pipecg timeout 300 $B --n 1024 --steps 1 --warmup 1 > $O/trace_run.json 2>> $O/err.txt
OHP_CG=pipecg timeout 300 $B --workload box3d_110 --steps 10 > $O/3d_pipecg.json 2>> $O/err.txt
OHP_CG=pipecg OHP_PF_GROUPS=3 timeout 300 $B --workload box3d_110 --steps 10 > $O/3d_pipecg_g3.json 2>> $O/err.txt
ls $O
