#!/bin/bash
# round-2 GPU call 22 (1 GPU): 3-D SpMV (15 non-zeros per row): sensitivity to the consumer configuration + one full ncu capture
set -x
O=gpurun_out/exp22; mkdir -p $O
B="python bench.py --no-cpu --secondary none --e2e-steps 0 --steps 10 --warmup 3 --sampler none --workload box3d_110"
timeout 300 $B > $O/3d_base.json 2>> $O/err.txt
for c in 1 2 3; do OHP_SPMV=$c timeout 300 $B > $O/3d_cfg$c.json 2>> $O/err.txt; done
OHP_IDX=32 timeout 300 $B > $O/3d_idx32.json 2>> $O/err.txt
OHP_L2PIN_MB=off timeout 300 $B > $O/3d_nohint.json 2>> $O/err.txt
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_spmv_tma -s 100 -c 2 -f -o $O/r02_spmv3d_prof $B --steps 1 --warmup 1 > $O/ncu.log 2>&1
ls $O
