#!/bin/bash
# round-2 GPU call 3 (1 GPU): whole GPU suite on the round-2 code, the default bench line, sampler perturbation after the
# pool allocator, and the Morton-ordered workload
set -x
O=gpurun_out/exp3; mkdir -p $O
timeout 1200 python -m pytest tests -m gpu -x -q > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $O/pytest_gpu.log
B="python bench.py --no-cpu --secondary none --e2e-steps 2 --steps 3 --warmup 3"
for s in "none 500" "smi 500" "nvml 500"; do set -- $s
  timeout 300 python bench.py --no-cpu --secondary none --workload box3d_110 --e2e-steps 3 --steps 30 --warmup 3 --sampler $1 --sampler-ms $2 > $O/sampler_3d_$1_$2.json 2>> $O/err.txt
  timeout 300 $B --sampler $1 --sampler-ms $2 > $O/sampler_full_$1_$2.json 2>> $O/err.txt
done
timeout 300 $B --workload box2d_2896_morton --sampler none > $O/morton.json 2>> $O/err.txt
timeout 600 python bench.py > $O/bench_default.json 2> $O/bench_default.err; echo "rc=$?" >> $O/bench_default.err
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > $O/bench_reference.json 2> $O/bench_reference.err
ls $O
