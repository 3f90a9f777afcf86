#!/bin/bash
# A/B the variants under variants/*.so against the in-tree libdrt.so on one GPU box: same command, same box.
# usage (inside gpurun): tools/ab_bench.sh [name ...]   -> one line per variant: name Msamples/s ms_per_step
cd "$(dirname "$0")/.."
names="$@"; [ -z "$names" ] && names=$(ls variants/*.so | sed 's#variants/libdrt_##; s#\.so##')
run() { python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-extras 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$1', round(d['value'],1), round(d['ms_per_step'],2), 'fp32', round((d.get('fp32_variant') or {}).get('value',0),1))"; }
run base
DRT_WAVE_FEAT=255 run generic_instantiation
for n in $names; do DRT_LIB=$PWD/variants/libdrt_$n.so run $n; done
run base
