#!/bin/bash
cd "$(dirname "$0")/.."
run() { (cd $1 && python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-extras 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$2', round(d['value'],1), round(d['ms_per_step'],2))"); }
run . HEAD
for c in 873aed3 d3a6d29 7b4d898 fa1f4b6 1349199; do run variants/wt/$c $c; done
run . HEAD
