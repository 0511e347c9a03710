#!/bin/bash
# round-2 GPU call 5 (1 GPU): GPU suite with the pipelined recurrence, then pipecg vs the two-launch single-reduction loop
# at the 8-, 4-, 2- and 1-GPU shares of config 4, with a device-side timeline of pipecg at the 8-GPU share
set -x
O=gpurun_out/exp5; mkdir -p $O
timeout 1200 python -m pytest tests -m gpu -q -x > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $O/pytest_gpu.log
B="python bench.py --no-cpu --secondary none --e2e-steps 1 --steps 3 --warmup 3 --sampler none"
for n in 1024 1448 2048; do
  for cg in cgcg pipecg; do
    OHP_CG=$cg timeout 300 $B --n $n > $O/n${n}_$cg.json 2>> $O/err.txt
  done
done
OHP_CG=pipecg timeout 300 $B > $O/full_pipecg.json 2>> $O/err.txt
OHP_CG=pipecg OHP_TRACE_CG=$O/trace_n1024_pipecg timeout 300 $B --n 1024 --steps 1 --warmup 1 > $O/trace_run.json 2>> $O/err.txt
OHP_CG=pipecg timeout 300 $B --workload box3d_110 --steps 10 > $O/3d_pipecg.json 2>> $O/err.txt
ls $O
