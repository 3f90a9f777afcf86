#!/bin/bash
# A/B of the non-headline configurations: tools/ab_configs.sh <variant name> [configs...]
cd "$(dirname "$0")/.."
v=$1; shift
cfgs="$@"; [ -z "$cfgs" ] && cfgs="c4 c5"
echo "== in-tree"; python tools/bench_configs.py $cfgs 2>/dev/null | python -c "import sys,json; [print(json.loads(l)['config'][:40], round(json.loads(l)['ms'],2)) for l in sys.stdin]"
echo "== $v"; DRT_LIB=$PWD/variants/libdrt_$v.so python tools/bench_configs.py $cfgs 2>/dev/null | python -c "import sys,json; [print(json.loads(l)['config'][:40], round(json.loads(l)['ms'],2)) for l in sys.stdin]"
