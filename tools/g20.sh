tools/gpu_round.sh g20 "default" 1000000 "planar relativistic nonlinear multi"
for w in planar relativistic; do
MCS_SCHED_STATS=1 MCS_LIB=$PWD/montecarloscattering.jl_b200/libmcs_b200_counters.so python bench.py --steps 1 --warmup 0 --no-cpu-baseline --no-verify --workload $w --n-per-pcut 1000000 > gpurun_out/g20_counters_$w.json 2> gpurun_out/g20_counters_$w.err
grep "mcs\]" gpurun_out/g20_counters_$w.err | tail -2
done
tools/profile.sh 200000 7 g20_2e5
tools/profile.sh 1000000 7 g20_1e6
WORKLOAD=relativistic tools/profile.sh 1000000 20 g20_rel
