tools/gpu_round.sh g6 "default"
TEST_LIB=tail tools/gpu_round.sh g6tail "tail tailb128 tailT8 tailT24 tailW4k"
SKIP_TESTS=1 tools/gpu_round.sh g6rel "default tail" 1000000 "relativistic multi"
