#!/bin/bash
# The closing GPU visit of a round: full -m gpu suite, smoke(), the default bench line and the reference arm as the driver runs
# them, then the documentation rows (45-pcut ladder, all species of config 5, 1e7 particles per pcut).  Run under gpurun.
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/final_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/final_pytest.log
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
( time python bench.py ) > gpurun_out/final_bench_default.json 2> gpurun_out/final_bench_default.err; tail -1 gpurun_out/final_bench_default.json | cut -c1-3000; tail -3 gpurun_out/final_bench_default.err
( time python bench.py --impl reference ) > gpurun_out/final_bench_reference.json 2> gpurun_out/final_bench_reference.err; tail -1 gpurun_out/final_bench_reference.json | cut -c1-1200; tail -3 gpurun_out/final_bench_reference.err
python bench.py --steps 2 --warmup 1 --no-cpu-baseline --pcuts default > gpurun_out/final_bench_planar45.json 2>/dev/null; tail -1 gpurun_out/final_bench_planar45.json | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('planar 45-pcut ladder %.3e steps/s, %d pcuts run, %.0f ms' % (d['value'], d['config']['pcuts_run'], d['ms_per_step']))"
python bench.py --steps 2 --warmup 1 --no-cpu-baseline --workload multi --all-species > gpurun_out/final_bench_multi_all.json 2>/dev/null; tail -1 gpurun_out/final_bench_multi_all.json | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('multi all species %.3e steps/s' % d['value'], {k:'%.3e' % v['steps_per_s'] for k,v in d['config']['species'].items()})"
python bench.py --steps 1 --warmup 1 --no-cpu-baseline --n-per-pcut 10000000 --generate-in-library > gpurun_out/final_bench_planar_1e7.json 2>/dev/null; tail -1 gpurun_out/final_bench_planar_1e7.json | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('planar 1e7 %.3e steps/s %.1f s/iter' % (d['value'], d['ms_per_step']/1e3))"
