#!/bin/bash
# Multi-GPU visit (run under gpurun --gpus N): the in-process N-rank parity tests, then bench.py under torchrun at N ranks
# (its `verify` block compares the N-rank NCCL path with one rank before the timed loop).
#   tools/gpu_multi.sh TAG N [extra bench args]
TAG=${1:-m2}; N=${2:-2}; shift 2
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv,noheader > gpurun_out/${TAG}_gpus.txt 2>&1
timeout 900 python -m pytest tests/test_parity_gpu.py -m gpu -x -q -k "nccl" > gpurun_out/${TAG}_pytest.log 2>&1
echo "pytest rc=$?"; tail -3 gpurun_out/${TAG}_pytest.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 \
  bench.py --gpus $N --steps 2 --warmup 1 "$@" > gpurun_out/${TAG}_bench_n$N.json 2> gpurun_out/${TAG}_bench_n$N.err
echo "bench rc=$?"
tail -1 gpurun_out/${TAG}_bench_n$N.json | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('N=%d value %.3e e2e %.3e ms/step %.0f' % (d['n_gpus'], d['value'], d['e2e']['value'], d['ms_per_step']))
print('verify', json.dumps(d.get('verify'))[:900])
print('per-rank kernel ms', d.get('per_rank_kernel_ms_per_step'))"
