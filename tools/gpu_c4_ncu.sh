#!/bin/bash
cd "$(dirname "$0")/.."
BAND_CONFIG=c4 BAND_Y0=300 BAND_H=480 python tools/profile_band.py 2>&1 | tail -1
BAND_CONFIG=c4 BAND_Y0=300 BAND_H=480 BAND_COUNT=1 python tools/profile_band.py 2>&1 | tail -1
BAND_CONFIG=c4 BAND_Y0=300 BAND_H=480 timeout 600 ncu --set full --clock-control none --import-source on -k regex:render_wave -s 2 -c 1 -o gpurun_out/r2_band_c4 -f python tools/profile_band.py > gpurun_out/ncu_band_c4.log 2>&1; echo "rc=$?"
