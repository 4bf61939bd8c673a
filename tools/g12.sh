tools/gpu_round.sh g12 "default" 1000000 "planar relativistic"
python tools/sched_dep.py relativistic 300000 45 > gpurun_out/g12_sched_rel.txt 2>&1; tail -8 gpurun_out/g12_sched_rel.txt
python tools/sched_dep.py nonlinear 300000 12 > gpurun_out/g12_sched_nl.txt 2>&1; tail -4 gpurun_out/g12_sched_nl.txt
