SKIP_TESTS=1 tools/gpu_round.sh g22 "default slim hyb" 1000000 "planar relativistic multi nonlinear"
