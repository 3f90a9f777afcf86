#!/bin/bash
# state check on one B200: full GPU suite, bench, launch list with counters, band capture, A/B against variants/*.so
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu.log
python bench.py --steps 5 --warmup 3 > gpurun_out/bench_head.json 2> gpurun_out/bench_head.err; echo "bench rc=$?"; cut -c1-300 gpurun_out/bench_head.json
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__inst_executed.sum,lts__t_sector_hit_rate.pct \
    --clock-control none --csv --log-file gpurun_out/launches_head.csv python bench.py --steps 1 --warmup 1 --no-extras --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1; echo "ncu launches rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:render_wave -s 2 -c 1 -o gpurun_out/r2_band_final -f python tools/profile_band.py > gpurun_out/ncu_band.log 2>&1; echo "ncu band rc=$?"; tail -1 gpurun_out/ncu_band.log
timeout 300 tools/ab_bench.sh 2>&1 | grep -v generic_instantiation
python tools/bench_configs.py c5 2>&1 | tail -1 | cut -c1-200
