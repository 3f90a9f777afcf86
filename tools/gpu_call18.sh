#!/bin/bash
cd "$(dirname "$0")/.."
timeout 900 python tools/gpu_fuzz.py 300 2>&1 | grep -i "mismatch"
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
timeout 300 tools/ab_bench.sh base 2>&1
