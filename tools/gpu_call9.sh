#!/bin/bash
# C5 after the Morton-code change: mesh parity tests + the full frame with counters
cd "$(dirname "$0")/.."
python -m pytest tests -m gpu -x -q -k "mesh or obj or multi" 2>&1 | tail -3
python tools/bench_configs.py c5 2>&1 | tail -1 | cut -c1-300
BAND_CONFIG=c5 BAND_COUNT=1 python tools/profile_band.py 2>&1 | tail -2
