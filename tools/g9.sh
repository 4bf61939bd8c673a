python tools/sched_dep.py relativistic 200000 8 > gpurun_out/g9_sched_rel.txt 2>&1; tail -12 gpurun_out/g9_sched_rel.txt
python tools/sched_dep.py planar 200000 4 > gpurun_out/g9_sched_planar.txt 2>&1; tail -6 gpurun_out/g9_sched_planar.txt
tools/gpu_round.sh g9 "default" 1000000 "planar"
MCS_DYNAMIC_QUEUE=1 SKIP_TESTS=1 tools/gpu_round.sh g9dyn "default" 1000000 "planar"
