#!/bin/bash
# round-2 GPU call 8 (1 GPU): L2 residency sweep at the 8-GPU share of config 4 (n=1024: 1.05 M rows, working set ~ L2 size):
# matrix tiles streamed with evict_first except OHP_L2PIN_MB megabytes kept with evict_last
set -x
O=gpurun_out/exp8; mkdir -p $O
B="python bench.py --no-cpu --secondary none --e2e-steps 0 --steps 3 --warmup 3 --sampler none"
for cg in cgcg pipecg; do
  for mb in 0 0.001 16 32 48 64 80; do
    OHP_CG=$cg OHP_L2PIN_MB=$mb timeout 300 $B --n 1024 > $O/n1024_${cg}_pin$mb.json 2>> $O/err.txt
  done
done
for mb in 0 0.001 32; do
  OHP_CG=cgcg OHP_L2PIN_MB=$mb timeout 300 $B --n 1448 > $O/n1448_cgcg_pin$mb.json 2>> $O/err.txt
done
ls $O
