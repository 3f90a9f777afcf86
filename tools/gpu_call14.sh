#!/bin/bash
cd "$(dirname "$0")/.."
timeout 600 ncu --set full --clock-control none --import-source on -k regex:render_wave -s 2 -c 1 -o gpurun_out/r2_band_chunk -f python tools/profile_band.py > gpurun_out/ncu_band.log 2>&1; echo "ncu band rc=$?"
DRT_LIB=$PWD/variants/libdrt_base.so timeout 600 ncu --set full --clock-control none --import-source on -k regex:render_wave -s 2 -c 1 -o gpurun_out/r2_band_base -f python tools/profile_band.py > gpurun_out/ncu_band2.log 2>&1; echo "ncu band rc=$?"
