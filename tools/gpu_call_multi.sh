#!/bin/bash
# multi-GPU call: the drt_render_multi test on all GPUs, then the bench line at N GPUs (frame-sharded value + tiles key)
cd "$(dirname "$0")/.."
N=${1:-2}
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/topo_$N.txt 2>&1
python -m pytest tests -m gpu -x -q -k "render_multi or sharding" > gpurun_out/pytest_multi_$N.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_multi_$N.log
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 3 --warmup 3 > gpurun_out/bench_${N}gpu.json 2> gpurun_out/bench_${N}gpu.err
echo "bench rc=$?"; tail -3 gpurun_out/bench_${N}gpu.err
python - <<PY
import json
d=json.loads(open('gpurun_out/bench_${N}gpu.json').read().strip().split('\n')[-1])
print('value', d['value'], 'e2e', d['e2e']['value'])
print(json.dumps(d.get('tiles'), indent=1))
PY
