#!/bin/bash
# Chebyshev time-to-solution on the bench workloads, then the whole 1-GPU pytest suite on the final code
mkdir -p gpurun_out
( timeout 120 python tools/cheb_measure.py gpurun_out/r02_chebyshev_time_to_solution.json 2>&1 | tail -30 ) > gpurun_out/r02_chebyshev_measure.log
tail -12 gpurun_out/r02_chebyshev_measure.log
( timeout 200 python -m pytest tests -q -m gpu 2>&1 | tail -30 ) > gpurun_out/r02_pytest_gpu_full_1gpu.log
tail -4 gpurun_out/r02_pytest_gpu_full_1gpu.log
