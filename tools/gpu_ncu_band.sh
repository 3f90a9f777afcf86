#!/bin/bash
# ncu --set full of one kernel on a band of a configuration: tools/gpu_ncu_band.sh <c2|c3|c4|c5> <kernel regex> [BAND_Y0 BAND_H]
#   -> gpurun_out/band_<config>.ncu-rep; summarise with tools/ncu_summary.py
cd "$(dirname "$0")/.."
cfg=${1:-c2}; k=${2:-render_wave}
[ -n "$3" ] && export BAND_Y0=$3; [ -n "$4" ] && export BAND_H=$4
BAND_CONFIG=$cfg python tools/profile_band.py 2>&1 | tail -1
BAND_CONFIG=$cfg timeout 600 ncu --set full --clock-control none --import-source on -k regex:$k -s 2 -c 1 -o gpurun_out/band_$cfg -f python tools/profile_band.py > gpurun_out/ncu_band_$cfg.log 2>&1; echo "ncu rc=$?"
