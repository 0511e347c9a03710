#!/bin/bash
# fused SpMV + Chebyshev update (k_spmv_cheb): parity tests, then time-to-solution again
mkdir -p gpurun_out
( timeout 120 python -m pytest tests/test_zz_next_rows_gpu.py -q -m gpu 2>&1 | tail -60 ) > gpurun_out/r02_pytest_next_rows_fused.log
tail -4 gpurun_out/r02_pytest_next_rows_fused.log
( PYTHONPATH=. timeout 100 python tools/cheb_measure.py gpurun_out/r02_chebyshev_time_to_solution_fused.json 2>&1 | tail -30 ) > gpurun_out/r02_chebyshev_measure_fused.log
cut -c1-330 gpurun_out/r02_chebyshev_measure_fused.log | tail -14
