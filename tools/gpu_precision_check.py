import sys
import os; sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from distraytracer_b200 import runtime, abi, scenes
scene, st = scenes.config2(1920, 1080, 64)
dev = runtime.DeviceScene(scene, 0)
for prec in (0, 1, 1, 0):
    st.precision = prec
    c = abi.Counters()
    dev.render_device(st, None, c); dev.render_device(st, None, c)
    print("precision", prec, round(c.kernel_ms, 1), "ms", round(132.7104 / c.kernel_ms * 1e3, 1), "Msamples/s", flush=True)
