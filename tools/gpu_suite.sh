#!/bin/bash
# the driver's GPU checks in one call: the full GPU suite (timed), smoke(), the bench line
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
t0=$(date +%s); timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -4; echo "pytest wall $(( $(date +%s) - t0 )) s"
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; echo "bench rc=$?"; cut -c1-160 gpurun_out/bench_default.json
