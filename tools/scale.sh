#!/bin/bash
# strong-scaling probe (run on a multi-GPU box): tools/scale.sh "2 4" [extra bench args]
NS=$1; shift
for n in $NS; do
  if [ "$n" = "1" ]; then L="python"; else L="python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29700"; fi
  timeout 400 $L bench.py --gpus $n --steps 2 --warmup 2 --e2e-steps 1 --no-cpu "$@" 2>gpurun_out/scale_err_$n.log | tail -1 > gpurun_out/scale_$n.json
  python -c "
import json; d=json.load(open('gpurun_out/scale_$n.json')); print($n, 'DOF/s %.3e'%d['value'], 'step_ms %.1f'%d['ms_per_step'], {k:round(v['ms'],4) for k,v in d['kernels'].items() if 'ms' in v}, 'e2e_ms %.0f'%d['e2e']['ms_per_step'])" || tail -20 gpurun_out/scale_err_$n.log
done
