tools/microbench > gpurun_out/g4_microbench.txt 2>&1
tools/gpu_round.sh g4 "default pf"
MCS_DET_TALLIES=0 SKIP_TESTS=1 tools/gpu_round.sh g4nodet "default"
