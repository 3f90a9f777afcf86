import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from distraytracer_b200 import runtime, abi, scenes
w, h, spp = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
scene, st = scenes.config2(w, h, spp)
dev = runtime.DeviceScene(scene, 0)
for collect in (0, 1):
    c = abi.Counters(); c.collect = collect
    try:
        dev.render_device(st, None, c)
        print(w, h, spp, "collect", collect, "ok", round(c.kernel_ms, 2), "ms rays", c.rays, flush=True)
    except Exception as e:
        print(w, h, spp, "collect", collect, "FAILED", e, flush=True); break
