#!/bin/bash
# round-2 GPU call 21 (1 GPU): small knobs of the two-launch pipelined loop at the 8-GPU share (n = 1024)
set -x
O=gpurun_out/exp21; mkdir -p $O
B="python bench.py --no-cpu --secondary none --e2e-steps 0 --steps 3 --warmup 3 --sampler none --n 1024"
OHP_CG=pipecg2 timeout 300 $B > $O/base.json 2>> $O/err.txt
for v in 2 3 4 6 8; do OHP_CG=pipecg2 OHP_VGRID_PER_SM=$v timeout 300 $B > $O/vgrid$v.json 2>> $O/err.txt; done
OHP_CG=pipecg2 OHP_NOPDL=1 timeout 300 $B > $O/nopdl.json 2>> $O/err.txt
OHP_CG=pipecg2 OHP_PIPE_RR=0 timeout 300 $B > $O/rr0.json 2>> $O/err.txt
OHP_CG=pipecg2 OHP_PIPE_RR=1000 timeout 300 $B > $O/rr1000.json 2>> $O/err.txt
ls $O
