tools/gpu_round.sh g15 "default prev" 1000000 "planar relativistic nonlinear"
for w in planar relativistic; do
MCS_SCHED_STATS=1 MCS_LIB=$PWD/montecarloscattering.jl_b200/libmcs_b200_counters.so python bench.py --steps 1 --warmup 0 --no-cpu-baseline --no-verify --workload $w --n-per-pcut 1000000 > gpurun_out/g15_counters_$w.json 2> gpurun_out/g15_counters_$w.err
grep "mcs\]" gpurun_out/g15_counters_$w.err | tail -4
done
python tools/sched_dep.py relativistic 300000 20 > gpurun_out/g15_sched_rel.txt 2>&1; tail -3 gpurun_out/g15_sched_rel.txt
