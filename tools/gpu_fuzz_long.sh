#!/bin/bash
# a long GPU-vs-oracle fuzz (mutated fixture scenes) + a fresh 1-GPU bench line; prints only the mismatching seeds
cd "$(dirname "$0")/.."
timeout 2400 python tools/gpu_fuzz.py ${1:-1000} 2>&1 | grep -i "mismatch"
python bench.py --steps 5 --warmup 3 > gpurun_out/bench_head.json 2> gpurun_out/bench_head.err; echo "bench rc=$?"; cut -c1-200 gpurun_out/bench_head.json
