#!/bin/bash
cd "$(dirname "$0")/.."
cat > /tmp/many.py <<'PY'
import sys, time; sys.path.insert(0, '.')
from distraytracer_b200 import scenes, runtime, abi
for nx, nz in ((24, 14), (48, 40), (96, 80)):
    sc, s = scenes.many_shapes(nx, nz, 1920, 1080, 16)
    dev = runtime.DeviceScene(sc, 0); cnt = abi.Counters()
    for i in range(3): dev.render_device(s, None, cnt)
    print("many_shapes", len(sc.prims), "prims 1080p 16spp: kernel", round(cnt.kernel_ms, 2), "ms", flush=True)
    dev.close()
PY
echo "== in-tree (tree over the geoms)"; python /tmp/many.py
echo "== base (linear slab filter from global memory)"; DRT_LIB=$PWD/variants/libdrt_base.so python /tmp/many.py
