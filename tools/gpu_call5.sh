#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -15 gpurun_out/pytest_gpu.log
python bench.py --no-cpu-baseline > gpurun_out/bench5.json 2> gpurun_out/bench5.err; echo "bench rc=$?"
python - <<PY
import json
d=json.loads(open('gpurun_out/bench5.json').read().strip().split('\n')[-1])
print('value', d['value'], 'e2e', d['e2e']['value'])
for k,v in d['configs'].items(): print(k, {a:b for a,b in v.items() if a in ('ms','frames_per_s','ms_per_frame','kernel_ms_per_frame','kernel_variant','error','Msamples_per_s')})
PY
