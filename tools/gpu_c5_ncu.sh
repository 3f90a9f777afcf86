#!/bin/bash
cd "$(dirname "$0")/.."
BAND_CONFIG=c5 BAND_COUNT=1 python tools/profile_band.py 2>&1 | tail -1
BAND_CONFIG=c5 timeout 600 ncu --set full --clock-control none --import-source on -k regex:render_wave -s 2 -c 1 -o gpurun_out/r2_c5_band_final -f python tools/profile_band.py > gpurun_out/c5_ncu.log 2>&1; echo "ncu rc=$?"
