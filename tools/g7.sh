tools/gpu_round.sh g7 "default" 1000000 "planar relativistic nonlinear"
tools/profile.sh 1000000 6 g7_1e6 > gpurun_out/g7_profile_1e6.log 2>&1
tools/profile.sh 200000 6 g7_2e5 > gpurun_out/g7_profile_2e5.log 2>&1
