#!/bin/bash
# round-2 GPU call 20 (1 GPU): which recurrence on ONE GPU between 0.4 M and 2.1 M rows (classic is the default there)
set -x
O=gpurun_out/exp20; mkdir -p $O
B="python bench.py --no-cpu --secondary none --e2e-steps 0 --steps 3 --warmup 3 --sampler none"
for cg in classic cgcg pipecg1; do
  OHP_CG=$cg timeout 300 $B --n 724 > $O/n724_$cg.json 2>> $O/err.txt
  OHP_CG=$cg timeout 300 $B --n 1024 > $O/n1024_$cg.json 2>> $O/err.txt
  OHP_CG=$cg timeout 300 $B --n 1448 > $O/n1448_$cg.json 2>> $O/err.txt
  OHP_CG=$cg timeout 300 $B --workload box3d_110 --steps 10 > $O/3d_$cg.json 2>> $O/err.txt
done
ls $O
