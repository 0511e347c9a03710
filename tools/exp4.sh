#!/bin/bash
# round-2 GPU call 4 (1 GPU): GPU suite, then the fused one-synchronisation CG kernel: parity (the CG tests with OHP_CG=fused)
# and speed against the two-launch loop at the 8-, 4-, 2- and 1-GPU shares of config 4
set -x
O=gpurun_out/exp4; mkdir -p $O
timeout 1200 python -m pytest tests -m gpu -q > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $O/pytest_gpu.log
OHP_CG=fused timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "cg or newton or solve or persistent or nonlinear or circle or monitors or morton" > $O/pytest_fused.log 2>&1; echo "pytest rc=$?" >> $O/pytest_fused.log
B="python bench.py --no-cpu --secondary none --e2e-steps 1 --steps 3 --warmup 3 --sampler none"
for n in 1024 1448 2048; do
  for cg in cgcg fused; do
    OHP_CG=$cg timeout 300 $B --n $n > $O/n${n}_$cg.json 2>> $O/err.txt
  done
  OHP_LIB=build/ab/libohp_fusedu2.so OHP_CG=fused timeout 300 $B --n $n > $O/n${n}_fusedu2.json 2>> $O/err.txt
done
OHP_CG=fused timeout 300 $B > $O/full_fused.json 2>> $O/err.txt
OHP_CG=fused timeout 300 $B --workload box3d_110 --steps 10 > $O/3d_fused.json 2>> $O/err.txt
timeout 300 $B --workload box3d_110 --steps 10 > $O/3d_default.json 2>> $O/err.txt
OHP_CG=fused OHP_TRACE_CG=$O/trace_n1024_fused timeout 300 $B --n 1024 --steps 1 --warmup 1 > $O/trace_run.json 2>> $O/err.txt
OHP_CG=fused OHP_L2PIN_MB=48 timeout 300 $B --n 1024 > $O/n1024_fused_pin48.json 2>> $O/err.txt
ls $O
