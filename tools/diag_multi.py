"""Where the wall clock of a single-frame render goes on one GPU: drt_render vs drt_render_multi with one handle."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
from distraytracer_b200 import runtime, abi, scenes
for tag, builder in (("C2", lambda: scenes.config2()), ("C3", scenes.config3), ("C5", scenes.config5)):
    scene, st = builder()
    dev = runtime.DeviceScene(scene, 0)
    frame = torch.empty((st.yRes, st.xRes, 3), dtype=torch.uint8).pin_memory().numpy()
    cnt = abi.Counters()
    for rep in range(3):
        t0 = time.perf_counter(); dev.render(st, out=frame, counters=cnt); t1 = time.perf_counter()
        a = (1e3 * (t1 - t0), cnt.kernel_ms)
        t0 = time.perf_counter(); _, c = runtime.render_multi([dev], st, out=frame, counters=True); t1 = time.perf_counter()
        b = (1e3 * (t1 - t0), c[0].kernel_ms, c[0].kernel_launches)
        t0 = time.perf_counter(); dev.render_device(st, None, cnt); t1 = time.perf_counter()
        d = (1e3 * (t1 - t0), cnt.kernel_ms)
        print(tag, "render wall/kernel %.2f %.2f | multi wall/kernel/launches %.2f %.2f %d | device-only wall/kernel %.2f %.2f" % (a + b + d), flush=True)
    dev.close()
