#!/bin/bash
# ncu evidence for the transport kernel (B200_PROFILING.md recipe). Run under gpurun; outputs land in gpurun_out/.
#   tools/profile.sh [n_per_pcut] [skip] [tag]
set -u
N=${1:-100000}
SKIP=${2:-5}
TAG=${3:-prof}
mkdir -p gpurun_out
CMD="python bench.py --steps 1 --warmup 0 --n-per-pcut $N --no-cpu-baseline --workload ${WORKLOAD:-planar}"
$CMD > gpurun_out/${TAG}_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${TAG}_launches.csv $CMD > gpurun_out/${TAG}_ncu_launches.log 2>&1
echo "launch list rc=$?"
$CMD > gpurun_out/${TAG}_plain2.log 2>&1 &&
ncu --set full --metrics lts__t_sectors_op_red.sum,lts__t_sectors_op_atom.sum --clock-control none --import-source on -k regex:transport_kernel -s $SKIP -c 1 -f -o gpurun_out/$TAG $CMD > gpurun_out/${TAG}_ncu_full.log 2>&1
echo "full capture rc=$?"
tail -1 gpurun_out/${TAG}_plain.log | cut -c1-400
