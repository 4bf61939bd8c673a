#!/bin/bash
# One GPU-box visit: parity tests on the default library, then time each build variant (run under gpurun).
#   tools/gpu_round.sh TAG "variant1 variant2 ..." [N_PER_PCUT] [WORKLOADS]
TAG=${1:-r02}
VARIANTS=${2:-"default"}
N=${3:-1000000}
WORKLOADS=${4:-"planar"}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv,noheader > gpurun_out/${TAG}_gpu.txt 2>&1
if [ -z "$SKIP_TESTS" ]; then
  # TEST_LIB=variant runs the parity tests on that build instead of the default library
  TL=""; [ -n "$TEST_LIB" ] && TL=$PWD/montecarloscattering.jl_b200/libmcs_b200_$TEST_LIB.so
  MCS_LIB=$TL timeout 1500 python -m pytest tests -m gpu -x -q ${PYTEST_K:+-k "$PYTEST_K"} > gpurun_out/${TAG}_pytest.log 2>&1
  echo "pytest rc=$?"; tail -5 gpurun_out/${TAG}_pytest.log
fi
for w in $WORKLOADS; do
for v in $VARIANTS; do
  s="_$v"; [ "$v" = default ] && s=""
  lib=$PWD/montecarloscattering.jl_b200/libmcs_b200$s.so
  [ -f "$lib" ] || { echo "missing $lib"; continue; }
  MCS_LIB=$lib timeout 400 python bench.py --steps 2 --warmup 1 --no-cpu-baseline --workload $w --n-per-pcut $N > gpurun_out/${TAG}_bench_${w}_$v.json 2> gpurun_out/${TAG}_bench_${w}_$v.err
  tail -1 gpurun_out/${TAG}_bench_${w}_$v.json | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$w variant $v', '%.3e steps/s' % d['value'], '%.0f ms' % d['ms_per_step'], 'frac %.4f' % d['roofline']['frac'], 'steps/iter', d['config']['steps_per_iteration'])" 2>&1 | tail -1
done
done
