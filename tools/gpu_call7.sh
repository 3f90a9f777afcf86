#!/bin/bash
cd "$(dirname "$0")/.."
python -m pytest tests -m gpu -x -q -k "mesh or obj" 2>&1 | tail -3
echo "== check build (pre-test rejected => exact rejects), full C5 frame"; DRT_LIB=$PWD/variants/libdrt_tricheck.so python tools/bench_configs.py c5 2>&1 | tail -1 | cut -c1-200
echo "== no pre-test"; DRT_LIB=$PWD/variants/libdrt_nopretest.so python tools/bench_configs.py c5 2>&1 | tail -1 | cut -c1-200
echo "== pre-test"; python tools/bench_configs.py c5 2>&1 | tail -1 | cut -c1-200
