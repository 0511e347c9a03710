#!/bin/bash
# round-2 GPU call 19 (8 GPUs): the driver's own command at N = 8 (library defaults, secondary 3-D workload, clock sampler)
set -x
O=gpurun_out/exp19; mkdir -p $O
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29700 bench.py --gpus 8 --steps 3 --warmup 3 > $O/n8_driver_command.json 2> $O/err.txt
echo "rc=$?"
ls $O
