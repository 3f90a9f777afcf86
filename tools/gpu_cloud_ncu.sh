#!/bin/bash
cd "$(dirname "$0")/.."
BAND_CONFIG=c3 BAND_Y0=400 BAND_H=64 timeout 600 ncu --set full --clock-control none --import-source on -k regex:cloud_corners -s 2 -c 1 -o gpurun_out/r2_band_c3_cloud -f python tools/profile_band.py > gpurun_out/ncu_band_c3.log 2>&1; echo "rc=$?"
