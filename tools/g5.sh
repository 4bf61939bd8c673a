tools/gpu_round.sh g5 "default ic128 icpf ic112"
