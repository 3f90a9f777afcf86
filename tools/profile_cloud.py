import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from distraytracer_b200 import runtime, abi, scenes
scene, st = scenes.config3(512, 288, 4)
st.cloud_only = 1
st.eye[:] = [0.5, 1.5, 1]; st.up[:] = [0, 0, 1]; st.lookingAt[:] = [0.5, -1, 1]
dev = runtime.DeviceScene(scene, 0)
cnt = abi.Counters()
for i in range(3):
    dev.render_device(st, None, cnt)
    print(f"cloud_only 512x288: kernel {cnt.kernel_ms:.3f} ms, {512*288/cnt.kernel_ms/1e3:.3f} Mcorners/s", flush=True)
