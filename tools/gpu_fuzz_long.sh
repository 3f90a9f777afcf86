#!/bin/bash
# long GPU-vs-oracle fuzz runs; prints only the mismatching seeds.  usage: tools/gpu_fuzz_long.sh "<n> <mode>" ["<n> <mode>" ...]
cd "$(dirname "$0")/.."
for spec in "$@"; do
  echo "== $spec"; timeout 2400 python tools/gpu_fuzz.py $spec 2>&1 | grep -i "mismatch\|error\|Traceback" | head -30
done
