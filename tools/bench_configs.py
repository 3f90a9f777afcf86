"""Device-timed throughput of every BASELINE.json configuration on one GPU (diagnostic table;
bench.py is the contractual line for configs[1])."""
import os, sys, time, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
from distraytracer_b200 import runtime, abi, scenes

def run(tag, scene, st, reps=2):
    t0 = time.time(); dev = runtime.DeviceScene(scene, 0); tb = time.time() - t0
    cnt = abi.Counters()
    spp = int(np.sqrt(st.antialias_samples)) ** 2
    n = st.xRes * st.yRes * spp
    ms = []
    for _ in range(reps + 1):
        dev.render_device(st, None, cnt); ms.append(cnt.kernel_ms)
    best = min(ms[1:])
    print(json.dumps({"config": tag, "res": [st.xRes, st.yRes], "spp": spp, "samples": n, "ms": best,
                      "Msamples_per_s": n / best / 1e3, "frames_per_s": 1e3 / best, "scene_create_s": tb,
                      "launches": cnt.kernel_launches}), flush=True)
    dev.close()

which = sys.argv[1:] or ["c1", "c2", "c3", "c4", "c5"]
if "c1" in which: run("C1 checkertexture 640x480 1spp", *scenes.config1())
if "c2" in which: run("C2 1080p 64spp all effects", *scenes.config2())
if "c3" in which: run("C3 Oren-Nayar + cloud background 1080p 256spp", *scenes.config3())
if "c4" in which: run("C4 mocap frame 30 velocity blur 1080p 16spp", *scenes.config4_frame(30))
if "c5" in which: run("C5 999698-triangle terrain 4K 64spp", *scenes.config5())
