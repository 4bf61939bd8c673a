#!/bin/bash
# time (and parity-check) each build variant of the transport kernel on the bench workload (run under gpurun)
#   tools/sweep_variants.sh N "suffix1 suffix2 ..."   ("" = the default library)
N=${1:-1000000}
VARIANTS=${2:-"default"}
for v in $VARIANTS; do
  s="_$v"; [ "$v" = default ] && s=""
  lib=$PWD/montecarloscattering.jl_b200/libmcs_b200$s.so
  [ -f "$lib" ] || { echo "missing $lib"; continue; }
  MCS_LIB=$lib timeout 300 python tools/quick_parity.py 2>&1 | grep -E "ALL OK|MISMATCH|Error" | head -3
  MCS_LIB=$lib timeout 300 python bench.py --steps 2 --warmup 1 --no-cpu-baseline --workload ${WORKLOAD:-planar} --n-per-pcut $N 2>&1 | tail -1 | \
    python -c "import json,sys; d=json.loads(sys.stdin.read()); print('variant $v', '%.3e steps/s' % d['value'], '%.0f ms' % d['ms_per_step'], 'frac %.4f' % d['roofline']['frac'])"
done
