#!/bin/bash
# round-2 GPU call 13 (1 GPU): dynamically scheduled SpMV in the two-launch pipelined CG: parity, then speed vs the static partition
set -x
O=gpurun_out/exp13; mkdir -p $O
timeout 1200 python -m pytest tests/test_gpu_parity.py tests/test_facade.py -m gpu -q > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $O/pytest_gpu.log
OHP_CG=pipecg2 timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "cg or newton or solve or nonlinear or circle or monitors or morton or full_size" > $O/pytest_pipecg2.log 2>&1; echo "pytest rc=$?" >> $O/pytest_pipecg2.log
B="python bench.py --no-cpu --secondary none --e2e-steps 0 --steps 3 --warmup 3 --sampler none"
for n in 1024 1448; do
  OHP_CG=pipecg2 OHP_NODYN=1 timeout 300 $B --n $n > $O/n${n}_pipecg2_static.json 2>> $O/err.txt
  for sh in 50 70 85; do
    OHP_CG=pipecg2 OHP_DYN_STATIC=$sh timeout 300 $B --n $n > $O/n${n}_pipecg2_dyn$sh.json 2>> $O/err.txt
  done
  OHP_CG=pipecg1 timeout 300 $B --n $n > $O/n${n}_pipecg1.json 2>> $O/err.txt
done
OHP_CG=pipecg2 OHP_PIPE_RR=500 timeout 300 $B --n 1024 > $O/n1024_pipecg2_rr500.json 2>> $O/err.txt
OHP_CG=pipecg2 OHP_TRACE_CG=$O/trace_n1024_pipecg2 timeout 300 $B --n 1024 --steps 1 --warmup 1 > $O/trace_run.json 2>> $O/err.txt
OHP_CG=pipecg1 OHP_TRACE_CG=$O/trace_n1024_pipecg1 timeout 300 $B --n 1024 --steps 1 --warmup 1 > $O/trace_run1.json 2>> $O/err.txt
timeout 300 $B > $O/full_default.json 2>> $O/err.txt
OHP_CG=pipecg2 timeout 300 $B --workload box3d_110 --steps 10 > $O/3d_pipecg2.json 2>> $O/err.txt
ls $O
