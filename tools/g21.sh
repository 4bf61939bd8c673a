SKIP_TESTS=1 tools/gpu_round.sh g21 "default ool slim" 1000000 "planar relativistic multi nonlinear"
