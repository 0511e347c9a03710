#!/bin/bash
# round-2 GPU call 17 (1 GPU): final records -- full GPU suite, default bench line, reference arm, ncu launch list + full
# captures of the dominant kernels (only after the same commands exited 0 without ncu)
set -x
O=gpurun_out/exp17; mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -q > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $O/pytest_gpu.log
timeout 900 python bench.py > $O/bench_n1.json 2> $O/bench_n1.err; echo "bench rc=$?"
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > $O/bench_reference.json 2> $O/bench_reference.err; echo "ref rc=$?"
CMD="python bench.py --steps 1 --warmup 1 --no-cpu --e2e-steps 1 --secondary none --sampler none"
timeout 300 $CMD > $O/prof_plain.log 2>&1 && \
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file $O/r02_launches.csv $CMD > $O/ncu1.log 2>&1
echo "launch list rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_spmv_tma -s 200 -c 2 -f -o $O/r02_spmv_prof $CMD > $O/ncu2.log 2>&1
echo "spmv capture rc=$?"
CMD2="python bench.py --steps 1 --warmup 1 --no-cpu --e2e-steps 0 --secondary none --sampler none --n 1024"
OHP_CG=pipecg1 timeout 300 $CMD2 > $O/prof_plain2.log 2>&1 && \
OHP_CG=pipecg1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_pipecg_fused -s 200 -c 2 -f -o $O/r02_pipefused_prof $CMD2 > $O/ncu3.log 2>&1
echo "fused capture rc=$?"
OHP_CG=pipecg2 timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_pipecg_vec -s 200 -c 2 -f -o $O/r02_pipevec_prof $CMD2 > $O/ncu4.log 2>&1
echo "vec capture rc=$?"
ls -la $O
