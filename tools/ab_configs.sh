#!/bin/bash
# A/B of the other BASELINE configurations: in-tree library vs variants/*.so (same box)
cd "$(dirname "$0")/.."
echo "== in-tree"; python tools/bench_configs.py c1 c3 c4 c5 2>&1 | grep '"config"' | python -c "import sys,json; [print('  ', d['config'][:40], round(d['ms'],2)) for d in map(json.loads, sys.stdin)]"
for v in variants/*.so; do echo "== $v"; DRT_LIB=$PWD/$v python tools/bench_configs.py c1 c3 c4 c5 2>&1 | grep '"config"' | python -c "import sys,json; [print('  ', d['config'][:40], round(d['ms'],2)) for d in map(json.loads, sys.stdin)]"; done
