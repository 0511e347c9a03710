#!/bin/bash
# round-2 final GPU call (6 GPU-minutes left): the new 8(f)-rank-4 tests first, then a subset of the hot-path parity tests and smoke()
mkdir -p gpurun_out
( timeout 200 python -m pytest tests/test_zz_next_rows_gpu.py -q -m gpu 2>&1 | tail -120 ) > gpurun_out/r02_pytest_next_rows.log
echo "next_rows rc=$?" >> gpurun_out/r02_pytest_next_rows.log
tail -5 gpurun_out/r02_pytest_next_rows.log
( timeout 150 python -m pytest tests/test_gpu_parity.py tests/test_facade.py -x -q -m gpu -k "not full_size" 2>&1 | tail -40 ) > gpurun_out/r02_pytest_parity_after_next_rows.log
tail -3 gpurun_out/r02_pytest_parity_after_next_rows.log
( timeout 60 python -c "import __graft_entry__ as g; g.smoke(); print('SMOKE_OK')" 2>&1 | tail -5 ) > gpurun_out/r02_smoke_after_next_rows.log
tail -2 gpurun_out/r02_smoke_after_next_rows.log
