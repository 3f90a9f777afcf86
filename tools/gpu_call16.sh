#!/bin/bash
cd "$(dirname "$0")/.."
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -6
python tools/bench_configs.py c5 2>&1 | tail -1 | cut -c1-200
python - <<'PY'
import sys, time; sys.path.insert(0, '.')
from distraytracer_b200 import scenes, runtime, abi
sc, s = scenes.many_shapes(48, 40, 1920, 1080, 16)     # 1921 shapes at 1080p 16 spp
dev = runtime.DeviceScene(sc, 0); cnt = abi.Counters()
for i in range(3): dev.render_device(s, None, cnt)
print("many_shapes", len(sc.prims), "prims 1080p 16spp: kernel", round(cnt.kernel_ms, 2), "ms, variant", cnt.kernel_variant if hasattr(cnt, 'kernel_variant') else '?')
PY
timeout 300 tools/ab_bench.sh 2>&1 | grep -v generic_instantiation
