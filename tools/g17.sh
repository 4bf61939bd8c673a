tools/gpu_round.sh g17 "default ssq" 1000000 "planar relativistic nonlinear"
