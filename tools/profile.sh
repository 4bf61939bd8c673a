#!/bin/bash
# ncu evidence for the transport kernel (B200_PROFILING.md recipe). Run under gpurun; outputs land in gpurun_out/.
#   tools/profile.sh [n_per_pcut]
set -u
N=${1:-100000}
mkdir -p gpurun_out
CMD="python bench.py --steps 1 --warmup 0 --n-per-pcut $N --no-cpu-baseline"
$CMD > gpurun_out/plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_launches.log 2>&1
echo "launch list rc=$?"
$CMD > gpurun_out/plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:transport_kernel -s 5 -c 2 -f -o gpurun_out/prof $CMD > gpurun_out/ncu_full.log 2>&1
echo "full capture rc=$?"
tail -2 gpurun_out/plain.log
