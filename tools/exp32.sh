#!/bin/bash
# the round's last GPU seconds: one ncu capture of k_spmv_cheb (4th launch: warm), sections for the roofline summary
mkdir -p gpurun_out
PYTHONPATH=. timeout 28 ncu --clock-control none --section SpeedOfLight --section MemoryWorkloadAnalysis --section Occupancy --section LaunchStats --section WarpStateStats \
  --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum -k regex:k_spmv_cheb --launch-skip 3 --launch-count 1 \
  -o gpurun_out/r02_spmv_cheb python tools/cheb_one.py > gpurun_out/r02_spmv_cheb_ncu.log 2>&1
echo "rc=$?" >> gpurun_out/r02_spmv_cheb_ncu.log
tail -5 gpurun_out/r02_spmv_cheb_ncu.log
