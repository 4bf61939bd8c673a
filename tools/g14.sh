SKIP_TESTS=1 tools/gpu_round.sh g14 "default" 1000000 "planar"
timeout 900 python -m pytest tests -m gpu -x -q -k "thermo" > gpurun_out/g14_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/g14_pytest.log
for w in planar relativistic; do
MCS_SCHED_STATS=1 MCS_LIB=$PWD/montecarloscattering.jl_b200/libmcs_b200_counters.so python bench.py --steps 1 --warmup 0 --no-cpu-baseline --no-verify --workload $w --n-per-pcut 1000000 > gpurun_out/g14_counters_$w.json 2> gpurun_out/g14_counters_$w.err
grep "mcs\]" gpurun_out/g14_counters_$w.err | tail -4
done
tools/profile.sh 200000 7 g14_2e5
tools/profile.sh 1000000 7 g14_1e6
