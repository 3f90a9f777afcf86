#!/bin/bash
# the value-noise background: golden + cloud tests, a sky fuzz, the C3 band (render + cloud + resolve launches), config 3 -- in-tree and variants/*.so
cd "$(dirname "$0")/.."
timeout 600 python -m pytest tests -m gpu -x -q -k "cloud or perlin or sky or config3 or render_multi" 2>&1 | tail -2
timeout 900 python tools/gpu_fuzz.py ${1:-150} sky 2>&1 | grep -i "mismatch"
run() { BAND_CONFIG=c3 BAND_Y0=400 BAND_H=64 python tools/profile_band.py 2>&1 | tail -1; python tools/bench_configs.py c3 2>&1 | tail -1 | cut -c60-140; }
echo "== in-tree"; run
for v in variants/*.so; do [ -f "$v" ] || continue; echo "== $v"; DRT_LIB=$PWD/$v run; done
