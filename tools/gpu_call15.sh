#!/bin/bash
cd "$(dirname "$0")/.."
BAND_CONFIG=c3 BAND_Y0=400 BAND_H=64 python tools/profile_band.py 2>&1 | tail -2
BAND_CONFIG=c3 BAND_Y0=400 BAND_H=64 timeout 600 ncu --set full --clock-control none --import-source on -k regex:render_wave -s 2 -c 1 -o gpurun_out/r2_band_c3 -f python tools/profile_band.py > gpurun_out/ncu_band_c3.log 2>&1; echo "rc=$?"
BAND_CONFIG=c3 BAND_Y0=400 BAND_H=64 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_c3.csv python tools/profile_band.py > /dev/null 2>&1
grep -c render_wave gpurun_out/launches_c3.csv; tail -8 gpurun_out/launches_c3.csv | cut -c1-200
