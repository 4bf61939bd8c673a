SKIP_TESTS=1 tools/gpu_round.sh g11 "default bins1" 1000000 "planar"
python tools/sched_dep.py relativistic 300000 45 > gpurun_out/g11_sched_rel.txt 2>&1; tail -14 gpurun_out/g11_sched_rel.txt
