"""Small, representative slice of the bench workload for ncu (a 1920x48 band through the
doors / glass block at 64 spp): keeps profiler replay time in seconds, not minutes."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from distraytracer_b200 import runtime, abi, scenes
scene, st = scenes.config2(1920, 1080, 64)
st.precision = int(os.environ.get("DRT_PRECISION", "0"))
dev = runtime.DeviceScene(scene, 0)
y0 = int(os.environ.get("BAND_Y0", "470")); h = int(os.environ.get("BAND_H", "48"))
tile = abi.Tile(0, y0, 1920, h, 0)
cnt = abi.Counters()
for i in range(3):
    dev.render_device(st, tile, cnt)
    print(f"band y0={y0} h={h}: kernel {cnt.kernel_ms:.3f} ms, {1920*h*64/cnt.kernel_ms/1e3:.1f} Msamples/s", flush=True)
