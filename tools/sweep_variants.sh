#!/bin/bash
# time each build variant of the transport kernel on the bench workload (run under gpurun)
N=${1:-1000000}
for v in "" _16x64 _16x128 _8x128 _8x64; do
  lib=$PWD/montecarloscattering.jl_b200/libmcs_b200$v.so
  [ -f "$lib" ] || continue
  MCS_LIB=$lib timeout 300 python bench.py --steps 1 --warmup 1 --no-cpu-baseline --n-per-pcut $N 2>&1 | tail -1 | \
    python -c "import json,sys; d=json.loads(sys.stdin.read()); print('variant ${v:-default}', '%.3e steps/s' % d['value'], '%.0f ms' % d['ms_per_step'], 'frac %.4f' % d['roofline']['frac'])"
done
