#!/bin/bash
# Statistical parity at the north-star size on the bench line itself: the GPU run at 1e7 particles per pcut against R
# independent replicas of the C port on the host cores (per-spectrum chi-square p-values in cpu_baseline.stat_parity).
mkdir -p gpurun_out
for w in nonlinear planar; do
python bench.py --steps 1 --warmup 1 --workload $w --n-per-pcut 10000000 --generate-in-library --stat-parity 8 --cpu-sample 80000 > gpurun_out/statparity_${w}_1e7.json 2> gpurun_out/statparity_${w}_1e7.err
tail -1 gpurun_out/statparity_${w}_1e7.json | python -c "
import json,sys
d=json.loads(sys.stdin.read()); sp=d['cpu_baseline']['stat_parity']
print('$w', '%.3e steps/s' % d['value'], d['cpu_baseline']['sample'])
for k,v in sp['spectra'].items(): print('   ', k, json.dumps(v)[:200])"
done
