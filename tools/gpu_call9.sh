#!/bin/bash
# C5: mesh parity tests + the full frame, in-tree library and the variants under variants/
cd "$(dirname "$0")/.."
python -m pytest tests -m gpu -x -q -k "mesh or obj or multi" 2>&1 | tail -3
echo "== in-tree"; python tools/bench_configs.py c5 2>&1 | tail -1 | cut -c1-200
for v in variants/*.so; do echo "== $v"; DRT_LIB=$PWD/$v python tools/bench_configs.py c5 2>&1 | tail -1 | cut -c1-200; done
BAND_CONFIG=c5 BAND_COUNT=1 python tools/profile_band.py 2>&1 | tail -2
