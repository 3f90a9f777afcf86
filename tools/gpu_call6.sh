#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q -k "cloud or perlin or render_multi or config3 or baseline_configs" 2>&1 | tail -5
python tools/bench_configs.py c3 2>&1 | tail -2
python tools/profile_cloud.py 2>&1 | tail -3
