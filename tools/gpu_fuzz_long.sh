#!/bin/bash
# a long GPU-vs-oracle fuzz (mutated fixture scenes; with "cam" the camera moves too); prints only the mismatching seeds
cd "$(dirname "$0")/.."
timeout 2400 python tools/gpu_fuzz.py ${1:-1000} $2 2>&1 | grep -i "mismatch\|error\|Traceback" | head -40
