SKIP_TESTS=1 tools/gpu_round.sh g18 "default pt8 pt8w pt24 psp128 psp512" 1000000 "relativistic planar"
