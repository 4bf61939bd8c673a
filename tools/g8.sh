tools/gpu_round.sh g8 "default" 1000000 "planar relativistic nonlinear"
MCS_DYNAMIC_QUEUE=1 SKIP_TESTS=1 tools/gpu_round.sh g8dyn "default" 1000000 "planar relativistic"
python bench.py --steps 2 --warmup 1 --no-cpu-baseline --workload multi --all-species > gpurun_out/g8_multi_all.json 2> gpurun_out/g8_multi_all.err
tail -1 gpurun_out/g8_multi_all.json | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('multi all-species %.3e steps/s' % d['value']); print(json.dumps(d['config'].get('species')))"
tools/profile.sh 1000000 6 g8_1e6 > gpurun_out/g8_profile_1e6.log 2>&1
tools/profile.sh 200000 6 g8_2e5 > gpurun_out/g8_profile_2e5.log 2>&1
