#!/bin/bash
# launch list with DRAM counters of the bench command (in-tree library)
cd "$(dirname "$0")/.."
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__inst_executed.sum,lts__t_sector_hit_rate.pct \
    --clock-control none --csv --log-file gpurun_out/launches_new.csv python bench.py --steps 1 --warmup 1 --no-extras --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1; echo "ncu launches rc=$?"
