SKIP_TESTS=1 tools/gpu_round.sh g16 "default prev lean leanz ssearch qip ssq" 1000000 "planar relativistic"
