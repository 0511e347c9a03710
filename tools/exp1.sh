#!/bin/bash
# round-2 GPU call 1: full GPU test suite, the new bench line, and the first L2-residency / CG-variant sweep
set -x
O=gpurun_out/exp1; mkdir -p $O
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $O/smi.txt
timeout 900 python -m pytest tests -m gpu -x -q > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $O/pytest_gpu.log
timeout 600 python bench.py --steps 2 --warmup 3 --e2e-steps 2 > $O/bench_default.json 2> $O/bench_default.err; echo "rc=$?" >> $O/bench_default.err
B="python bench.py --no-cpu --secondary none --e2e-steps 1 --steps 3 --warmup 3"
# the 8-GPU share of config 4 on ONE GPU (N = 1.05 M rows): which CG variant, and does pinning the matrix in L2 pay?
for cg in cgcg classic persistent; do
  for pin in 0 48 80 100; do
    OHP_CG=$cg OHP_L2PIN_MB=$pin timeout 300 $B --n 1024 > $O/n1024_${cg}_pin${pin}.json 2>> $O/sweep.err
  done
done
# 4-GPU and 2-GPU shares
for n in 1448 2048; do for pin in 0 64 100; do
  OHP_L2PIN_MB=$pin timeout 300 $B --n $n > $O/n${n}_auto_pin${pin}.json 2>> $O/sweep.err
done; done
# full size
for pin in 0 32 64 96; do
  OHP_L2PIN_MB=$pin timeout 300 $B > $O/full_pin${pin}.json 2>> $O/sweep.err
done
# device-side timeline of the two-launch single-reduction loop at the 8-GPU share
OHP_CG=cgcg OHP_TRACE_CG=$O/trace_n1024_cgcg timeout 300 $B --n 1024 --steps 1 --warmup 1 > $O/trace_run.json 2>> $O/sweep.err
OHP_CG=classic OHP_TRACE_CG=$O/trace_n1024_classic timeout 300 $B --n 1024 --steps 1 --warmup 1 >> $O/trace_run.json 2>> $O/sweep.err
ls -la $O
