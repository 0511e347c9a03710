#!/bin/bash
# Chebyshev time-to-solution on the bench workloads (retry of exp26 with the repo root on the module path)
mkdir -p gpurun_out
( PYTHONPATH=. timeout 150 python tools/cheb_measure.py gpurun_out/r02_chebyshev_time_to_solution.json 2>&1 | tail -30 ) > gpurun_out/r02_chebyshev_measure.log
tail -14 gpurun_out/r02_chebyshev_measure.log
