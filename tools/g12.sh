tools/gpu_round.sh g12 "default" 1000000 "planar relativistic"
python tools/sched_dep.py relativistic 300000 45 > gpurun_out/g12_sched_rel.txt 2>&1; tail -8 gpurun_out/g12_sched_rel.txt
python tools/sched_dep.py nonlinear 300000 12 > gpurun_out/g12_sched_nl.txt 2>&1; tail -4 gpurun_out/g12_sched_nl.txt
TEST_LIB=sacc PYTEST_K="deterministic or exact or dynamic or per_particle_parity" tools/gpu_round.sh g12sacc "sacc" 1000000 "planar"
MCS_DYNAMIC_QUEUE=1 SKIP_TESTS=1 tools/gpu_round.sh g12dyn "default sacc" 1000000 "planar relativistic"
