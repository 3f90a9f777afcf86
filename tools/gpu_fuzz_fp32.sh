#!/bin/bash
# the single-precision build on the fuzz generators (sanity: same pictures up to rounding-fragile pixels)
cd "$(dirname "$0")/.."
export DRT_FUZZ_PRECISION=1
for spec in "150 big" "150 meshes" "100 meshmotion" "150 scenes"; do echo "== fp32 $spec"; timeout 600 python tools/gpu_fuzz.py $spec 2>&1 | grep -i "mismatch\|error\|Traceback" | cut -c1-160 | tail -6; done
