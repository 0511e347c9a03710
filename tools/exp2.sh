#!/bin/bash
# round-2 GPU call 2: is the SpMV slower because of the box or because of the round-2 kernel changes?  A/B of library
# builds on ONE box, a device-copy calibration, and how much the clock sampler perturbs the step time.
set -x
O=gpurun_out/exp2; mkdir -p $O
python - > $O/copy_bw.txt 2>&1 <<'PY'
import torch, time
a=torch.empty(1<<30, dtype=torch.bfloat16, device='cuda'); b=torch.empty_like(a)
for _ in range(3): b.copy_(a)
best=1e9
for _ in range(10):
    e0=torch.cuda.Event(enable_timing=True); e1=torch.cuda.Event(enable_timing=True)
    e0.record(); b.copy_(a); e1.record(); torch.cuda.synchronize(); best=min(best,e0.elapsed_time(e1))
print("copy GB/s", 2*a.numel()*2/best/1e6)
PY
B="python bench.py --no-cpu --secondary none --e2e-steps 1 --steps 3 --warmup 3"
for lib in r01 none noesc noprefetch nocoh cur; do
  L=build/ab/libohp_$lib.so; [ $lib = cur ] && L=omegahpetsc_b200/lib/libohp.so
  OHP_LIB=$L timeout 300 $B --sampler none > $O/full_$lib.json 2>> $O/err.txt
  OHP_LIB=$L OHP_CG=cgcg timeout 300 $B --sampler none --n 1024 > $O/n1024_$lib.json 2>> $O/err.txt
done
# sampler perturbation (current library, full size and the 3-D box)
for s in "none 500" "smi 200" "smi 1000" "nvml 200" "nvml 1000"; do set -- $s
  timeout 300 $B --sampler $1 --sampler-ms $2 > $O/sampler_full_$1_$2.json 2>> $O/err.txt
  timeout 300 python bench.py --no-cpu --secondary none --workload box3d_110 --e2e-steps 1 --steps 30 --warmup 3 --sampler $1 --sampler-ms $2 > $O/sampler_3d_$1_$2.json 2>> $O/err.txt
done
ls $O
