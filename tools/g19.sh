SKIP_TESTS=1 tools/gpu_round.sh g19 "default pt20 pt24 pt24w4 pt24w8 pt28" 1000000 "relativistic planar nonlinear"
SKIP_TESTS=1 tools/gpu_round.sh g19m "default pt24 pt24w8" 1000000 "multi"
