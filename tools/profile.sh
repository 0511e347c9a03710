#!/bin/bash
# Run ON THE GPU BOX (under gpurun): launch list + one full ncu capture of the named kernel.
# usage: tools/profile.sh <tag> <kernel-regex> [skip] [bench args...]
set -u
TAG=$1; KREGEX=$2; SKIP=${3:-200}; shift 3 || true
CMD="python bench.py --steps 1 --warmup 1 --no-cpu --e2e-steps 1 $*"
mkdir -p gpurun_out
$CMD > gpurun_out/${TAG}_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/${TAG}_launches.csv $CMD > gpurun_out/${TAG}_ncu1.log 2>&1
echo "launch list rc=$?"
ncu --set full --clock-control none --import-source on -k regex:${KREGEX} -s ${SKIP} -c 2 -f -o gpurun_out/${TAG}_prof $CMD > gpurun_out/${TAG}_ncu2.log 2>&1
echo "full capture rc=$?"
tail -2 gpurun_out/${TAG}_plain.log | cut -c1-400
