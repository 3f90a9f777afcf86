"""Repeat one render N times: kernel time per run + hash of the u8 frame (must not change run to run)."""
import hashlib, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
from distraytracer_b200 import runtime, abi, scenes
w, h, spp, n = (int(x) for x in sys.argv[1:5])
collect = int(sys.argv[5]) if len(sys.argv) > 5 else 0
scene, st = scenes.config2(w, h, spp)
dev = runtime.DeviceScene(scene, 0)
seen = {}
for i in range(n):
    c = abi.Counters(); c.collect = collect
    img = dev.render(st, None, counters=c)
    hsh = hashlib.sha1(np.ascontiguousarray(img).tobytes()).hexdigest()[:12]
    seen[hsh] = seen.get(hsh, 0) + 1
    print(i, round(c.kernel_ms, 1), "ms", hsh, c.rays if collect else "", flush=True)
print("distinct frames:", len(seen), seen)
