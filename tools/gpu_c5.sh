#!/bin/bash
# config 5 and the mesh tests (in-tree library and variants/*.so)
cd "$(dirname "$0")/.."
timeout 600 python -m pytest tests -m gpu -x -q -k "mesh or obj" 2>&1 | tail -2
echo "== in-tree"; for i in 1 2; do python tools/bench_configs.py c5 2>&1 | tail -1 | cut -c60-130; done
for v in variants/*.so; do [ -f "$v" ] || continue; echo "== $v"; DRT_LIB=$PWD/$v python tools/bench_configs.py c5 2>&1 | tail -1 | cut -c60-130; done
