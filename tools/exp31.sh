#!/bin/bash
# last seconds of the round's GPU budget: the recorded-values test and the CG parity cases (Jacobi work-vector allocation change)
mkdir -p gpurun_out
( timeout 25 python -m pytest tests/test_zz_next_rows_gpu.py tests/test_gpu_parity.py -q -m gpu -k "recorded or cg_matches_oracle or persistent_cg" 2>&1 | tail -15 ) > gpurun_out/r02_pytest_last.log
tail -4 gpurun_out/r02_pytest_last.log
