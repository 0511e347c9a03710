#!/bin/bash
# round-2 GPU call 10 (1 GPU): one-launch-per-iteration pipelined CG -- parity tests, then speed at the 8/4/2/1-GPU shares
set -x
O=gpurun_out/exp10; mkdir -p $O
timeout 1200 python -m pytest tests -m gpu -q > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $O/pytest_gpu.log
B="python bench.py --no-cpu --secondary none --e2e-steps 0 --steps 3 --warmup 3 --sampler none"
for n in 1024 1448 2048; do
  for cg in cgcg pipecg2 pipecg; do
    OHP_CG=$cg timeout 300 $B --n $n > $O/n${n}_$cg.json 2>> $O/err.txt
  done
done
timeout 300 $B > $O/full_default.json 2>> $O/err.txt
OHP_CG=pipecg timeout 300 $B > $O/full_pipecg.json 2>> $O/err.txt
OHP_CG=pipecg OHP_TRACE_CG=$O/trace_n1024_pipecg timeout 300 $B --n 1024 --steps 1 --warmup 1 > $O/trace_run.json 2>> $O/err.txt
OHP_CG=pipecg timeout 300 $B --workload box3d_110 --steps 10 > $O/3d_pipecg.json 2>> $O/err.txt
timeout 300 $B --workload box3d_110 --steps 10 > $O/3d_default.json 2>> $O/err.txt
ls $O
