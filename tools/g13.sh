tools/gpu_round.sh g13 "default" 1000000 "planar nonlinear"
