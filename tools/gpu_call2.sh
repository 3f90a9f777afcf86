#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
tools/ab_bench.sh > gpurun_out/ab2.txt 2>&1; cat gpurun_out/ab2.txt
python tools/profile_band.py > gpurun_out/band.log 2>&1; tail -2 gpurun_out/band.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:render_wave -c 1 -s 2 -o gpurun_out/r2_band_specialised -f python tools/profile_band.py > gpurun_out/ncu_band.log 2>&1
echo "ncu rc=$?"; ls -la gpurun_out/*.ncu-rep
