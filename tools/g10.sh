python tools/sched_dep.py relativistic 200000 8 > gpurun_out/g10_sched_rel.txt 2>&1; tail -14 gpurun_out/g10_sched_rel.txt
python tools/sched_dep.py planar 200000 4 > gpurun_out/g10_sched_planar.txt 2>&1; tail -6 gpurun_out/g10_sched_planar.txt
